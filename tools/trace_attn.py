#!/usr/bin/env python
"""Dump the clock64 event trace of CTA 0 of the TC2 attention kernel (c3 layer shape) as per-tile deltas."""
import ctypes as C, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from rmem_b200 import _capi, ops as K

dev = torch.device("cuda:0"); lib = _capi.load(); OP = _capi.op_dtype()
T, HW = 8, 1674
g = torch.Generator().manual_seed(0)
q = torch.randn(HW, 128, generator=g).to(dev).to(OP)
k = torch.randn(T, HW, 128, generator=g).to(dev); v = torch.randn(T, HW, 1024, generator=g).to(dev)
kb, vtb, HWp = K.build_bank(k, v, 9, list(range(T)))
pe_cur = torch.zeros(128, device=dev); pe_mem = torch.zeros(4, 128, device=dev)
for _ in range(3):
    K.long_attention(q, kb, vtb, list(range(T)), HW, pe_cur, pe_mem, impl=2)
tr = torch.zeros(256 * 16, dtype=torch.int64, device=dev)
_capi.check(lib.rmem_debug_attn_trace(C.c_void_p(tr.data_ptr())))
K.long_attention(q, kb, vtb, list(range(T)), HW, pe_cur, pe_mem, impl=2)
torch.cuda.synchronize()
_capi.check(lib.rmem_debug_attn_trace(C.c_void_p(0)))
t = tr.cpu().view(256, 16)
t0 = int(t[t > 0].min())
names = ["pfull_seen", "pv_issued", "s_waits_done", "s_issued", "sm_start", "sfull_seen", "max_done", "handoff_done", "exp_done", "p_arrived", "epi_start", "epi_end", "epi_exch", "epi_pvdone"]
print("tile " + " ".join(n.rjust(12) for n in names))
for j in range(90):
    if int(t[j].max()) == 0:
        break
    print(f"{j:4d} " + " ".join((str(int(x) - t0) if int(x) > 0 else "-").rjust(12) for x in t[j, :14]))

ct = t[100:248]
g0 = int(ct[:, 0][ct[:, 0] > 0].min())
print("per-CTA wall time (ns from first start): cta start end tiles segs smid cycles")
rows = [(c, int(ct[c, 0]) - g0, int(ct[c, 1]) - g0, int(ct[c, 2]), int(ct[c, 3]), int(ct[c, 4]), int(ct[c, 6] - ct[c, 5])) for c in range(148) if int(ct[c, 0]) > 0]
for r in rows[:6] + sorted(rows, key=lambda r: -r[2])[:10]:
    print("  ", r)
print("max end", max(r[2] for r in rows), "min end", min(r[2] for r in rows), "max start", max(r[1] for r in rows))
