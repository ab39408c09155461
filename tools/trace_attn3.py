#!/usr/bin/env python
"""Clock64 event trace of one cluster (leader + peer CTA) of the TC3 pair attention kernel at the c3 layer shape, printed
as cycles since the first event.  RMEM_TRACE_CTA selects the cluster."""
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
    K.long_attention(q, kb, vtb, list(range(T)), HW, impl=3, grid=(31, 54))
ROWS = 760
tr = torch.zeros(ROWS * 16, dtype=torch.int64, device=dev)
# the local-attention / tc2 trace hooks share the entry point and write rows < 248 of the same buffer; harmless here
_capi.check(lib.rmem_debug_attn_trace(C.c_void_p(tr.data_ptr())))
K.long_attention(q, kb, vtb, list(range(T)), HW, impl=3, grid=(31, 54))
torch.cuda.synchronize()
_capi.check(lib.rmem_debug_attn_trace(C.c_void_p(0)))
t = tr.cpu().view(ROWS, 16)
names = ["pfull_seen", "pv_issued", "s_waits_done", "s_issued", "sm_start", "sfull_seen", "max_done", "handoff_done",
         "exp_done", "p_arrived", "s_kfull_seen", "p_arr_q1", "p_arr_q2", "vfull_seen", "p_arr_q3"]
for rank in range(2):
    tt = t[rank * 256:(rank + 1) * 256]
    if int(tt.max()) == 0:
        continue
    t0 = int(tt[tt > 0].min())
    print(f"rank {rank} (cycles since first event); sub-tile j, S events on even j")
    print("   j " + " ".join(n.rjust(12) for n in names))
    for j in range(96):
        if int(tt[j].max()) == 0:
            break
        print(f"{j:4d} " + " ".join((str(int(x) - t0) if int(x) > 0 else "-").rjust(12) for x in tt[j, :15]))
ct = t[600:748]
g0 = int(ct[:, 0][ct[:, 0] > 0].min())
rows = [(c, int(ct[c, 0]) - g0, int(ct[c, 1]) - g0, int(ct[c, 2]), int(ct[c, 3]), int(ct[c, 4]), int(ct[c, 6] - ct[c, 5]))
        for c in range(148) if int(ct[c, 0]) > 0]
print("phase split of even (leader) CTAs, cycles from kernel entry: cta setup_done prev_kernel_done first_PV last_PV epi0 epi1 end")
for c in list(range(0, 12, 2)) + list(range(60, 72, 2)):
    if int(ct[c, 0]) > 0:
        b = int(ct[c, 5])
        print("  ", c, *[(int(ct[c, k]) - b if int(ct[c, k]) > 0 else "-") for k in (7, 8, 9, 10, 11, 12, 6)])
print("per-CTA wall time (ns from first start): cta start end groups segs smid cycles")
for r in rows[:6] + sorted(rows, key=lambda r: -r[2])[:10]:
    print("  ", r)
print("max end", max(r[2] for r in rows), "min end", min(r[2] for r in rows), "max start", max(r[1] for r in rows))
