import sys, os, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from rmem_b200 import _capi, ops as K
dev = torch.device("cuda:0"); OP = _capi.op_dtype()
T, HW = 8, 1674
for name, sharp, hot, pe in [("plain", 1.0, False, False), ("sharp2", 2.0, False, False), ("hot", 1.0, True, False),
                             ("sharp2+hot", 2.0, True, False), ("pe", 1.0, False, True), ("sharp6", 6.0, False, False)]:
    g = torch.Generator().manual_seed(1)
    q = (torch.randn(HW, 128, generator=g) * sharp).to(dev).to(OP)
    k = torch.randn(T, HW, 128, generator=g)
    if hot:
        k = k * torch.linspace(0.5, 1.5, HW).view(1, HW, 1)
    k = k.to(dev); v = torch.randn(T, HW, 1024, generator=g).to(dev)
    slots = list(range(T))
    kb, vtb, HWp = K.build_bank(k, v, T + 1, slots)
    kw = {}
    if pe:
        kw = dict(pe_cur=(torch.randn(128, generator=g) * 0.1).to(dev), mem_pos_emb=(torch.randn(4, 128, generator=g) * 0.5).to(dev))
    od, md = K.long_attention(q, kb, vtb, slots, HW, impl=0, **kw)
    ot, mt = K.long_attention(q, kb, vtb, slots, HW, impl=2, **kw)
    torch.cuda.synchronize()
    err = (ot.float() - od.float()).abs()
    bad = (err.max(dim=1).values > 0.05).nonzero().flatten()
    nq = (HW + 127) // 128
    tab = [[int(err[qi * 128:(qi + 1) * 128, c * 256:(c + 1) * 256].max() > 0.05) for c in range(4)] for qi in range(nq)]
    print(f"{name:12s} max err {float(err.max()):.3e} bad rows {bad.numel()} first {bad[:8].tolist()} finite {bool(torch.isfinite(ot.float()).all())}")
    if bad.numel():
        print("   per (q-tile, dv-chunk):", tab)
