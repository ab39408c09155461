#!/usr/bin/env python
"""clock64 event trace of CTA 0 of the tensor-core local (windowed) attention kernel at the c3 shape."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from rmem_b200 import _capi, ops as K
dev = torch.device("cuda:0"); lib = _capi.load(); OP = _capi.op_dtype()
h, w = 31, 54; HW = h * w
g = torch.Generator().manual_seed(0)
q = torch.randn(HW, 128, generator=g).to(dev).to(OP); k = torch.randn(HW, 128, generator=g).to(dev).to(OP)
v = torch.randn(HW, 1024, generator=g).to(dev).to(OP); gate = torch.randn(HW, 1024, generator=g).to(dev).to(OP)
rw = (torch.randn(225, 128, generator=g) * 0.1).to(dev); rb = (torch.randn(225, generator=g) * 0.1).to(dev)
for _ in range(5):
    K.local_attention(q, k, v, rw, rb, h, w, gate)
tr = torch.zeros(256 * 16, dtype=torch.int64, device=dev)
_capi.check(lib.rmem_debug_attn_trace(C.c_void_p(tr.data_ptr())))
K.local_attention(q, k, v, rw, rb, h, w, gate)
torch.cuda.synchronize()
_capi.check(lib.rmem_debug_attn_trace(C.c_void_p(0)))
t = tr.cpu().view(256, 16)
t0 = int(t[40, 0])
names = ["pfull_seen", "pv_issued", "s_waits_done", "s_issued", "sm_start", "sfull_seen", "mask_max", "handoff", "exp_done", "p_arrived", "vfull_seen"]
print("tile " + " ".join(n.rjust(12) for n in names))
for j in range(18):
    print(f"{j:4d} " + " ".join((str(int(x) - t0) if int(x) > 0 else "-").rjust(12) for x in t[j, :11]))
print("kernel: start 0, softmax loop end", int(t[40, 1]) - t0, "last PV done", int(t[40, 2]) - t0, "epilogue end", int(t[40, 3]) - t0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    K.local_attention(q, k, v, rw, rb, h, w, gate)
e1.record(); torch.cuda.synchronize()
print("op (rel GEMM + transpose + kernel) us:", e0.elapsed_time(e1) / 20 * 1e3)
