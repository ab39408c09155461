#!/usr/bin/env python
"""clock64 event trace of the first CTAs of one gemm_tc launch (per-CTA deltas in cycles)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from rmem_b200 import _capi, ops as K
dev = torch.device("cuda:0"); lib = _capi.load(); OP = _capi.op_dtype()
for (M, N, Kd) in [(25773, 256, 64), (6527, 128, 1152), (1674, 1024, 256), (1674, 256, 1024), (1674, 256, 2304)]:
    A = torch.randn(M, Kd, device=dev).to(OP); W = torch.randn(N, Kd, device=dev).to(OP); b = torch.randn(N, device=dev)
    for _ in range(3):
        K.gemm(A, W, b, act=K.ACT_RELU)
    tr = torch.zeros(64 * 8, dtype=torch.int64, device=dev)
    _capi.check(lib.rmem_debug_gemm_trace(C.c_void_p(tr.data_ptr())))
    K.gemm(A, W, b, act=K.ACT_RELU)
    torch.cuda.synchronize()
    _capi.check(lib.rmem_debug_gemm_trace(C.c_void_p(0)))
    t = tr.cpu().view(64, 8)
    print(f"M={M} N={N} K={Kd}: per-CTA cycles  [setup, first tile landed, last tile landed, accumulator ready, chunk0: tmem loaded, math done, stores issued]")
    for c in (0, 1, 2, 30, 55, 63):
        r = t[c]
        if int(r[0]) == 0: continue
        print("   cta", c, [int(r[k] - r[0]) for k in (1, 2, 3, 4, 5, 6, 7)])
