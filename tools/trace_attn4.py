#!/usr/bin/env python
"""Per-CTA phase times of the column attention kernel (long_attn_tc4_kernel) at the c3 layer shape: cycles from kernel entry
to set-up done, previous kernel done, pass-0 softmax done, pass 0..3 complete, last read-out stored, end."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from rmem_b200 import _capi, ops as K

dev = torch.device("cuda:0"); lib = _capi.load(); OP = _capi.op_dtype()
T, HW = 8, 1674
g = torch.Generator().manual_seed(0)
q = torch.randn(HW, 128, generator=g).to(dev).to(OP)
k = torch.randn(T, HW, 128, generator=g).to(dev); v = torch.randn(T, HW, 1024, generator=g).to(dev)
kb, vtb, HWp = K.build_bank(k, v, 9, list(range(T)))
for _ in range(3):
    K.long_attention(q, kb, vtb, list(range(T)), HW, impl=4, grid=(31, 54))
ROWS = 760
tr = torch.zeros(ROWS * 16, dtype=torch.int64, device=dev)
_capi.check(lib.rmem_debug_attn_trace(C.c_void_p(tr.data_ptr())))
K.long_attention(q, kb, vtb, list(range(T)), HW, impl=4, grid=(31, 54))
torch.cuda.synchronize()
_capi.check(lib.rmem_debug_attn_trace(C.c_void_p(0)))
print("fallback flag", K.last_attn_overflow)
ct = tr.cpu().view(ROWS, 16)[600:748]
g0 = int(ct[:, 0][ct[:, 0] > 0].min())
print("cta n  start_ns end_ns | cycles from entry: setup prev_done softmax0_done pass0 pass1 pass2 pass3 stored end")
for c in list(range(0, 8)) + list(range(80, 88)) + [146, 147]:
    if int(ct[c, 0]) == 0:
        continue
    b = int(ct[c, 5])
    print(f"{c:4d} {int(ct[c, 2]):2d} {int(ct[c, 0]) - g0:7d} {int(ct[c, 1]) - g0:7d} |",
          *[(int(ct[c, k]) - b if int(ct[c, k]) > 0 else "-") for k in (7, 8, 13, 9, 10, 11, 12, 14, 6)])
ends = [int(ct[c, 1]) - g0 for c in range(148) if int(ct[c, 0]) > 0]
print("max end ns", max(ends), "min end ns", min(ends))
