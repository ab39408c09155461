#!/usr/bin/env python
"""Micro-benchmark: tcgen05 GEMM/conv kernel vs the legacy mma.sync kernel on the engine's shapes (CUDA events)."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rmem_b200 import _capi, ops as K  # noqa: E402


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    dev = torch.device("cuda:0")
    OP = _capi.op_dtype()
    lib = _capi.load()
    lin = [(25773, 64, 64), (25773, 256, 64), (25773, 64, 256), (6527, 128, 512), (6527, 512, 128), (1674, 256, 1024),
           (1674, 1024, 256), (1674, 640, 256), (1674, 512, 1024), (1674, 512, 256), (1674, 128, 512)]
    print("linear M N K : tc us | legacy us | TF/s tc")
    for M, N, Kd in lin:
        A = torch.randn(M, Kd, device=dev).to(OP)
        W = torch.randn(N, Kd, device=dev).to(OP)
        b = torch.randn(N, device=dev)
        t = []
        for impl in (0, 1):
            lib.rmem_set_gemm_impl(impl)
            t.append(timeit(lambda: K.gemm(A, W, b, act=K.ACT_RELU)))
        lib.rmem_set_gemm_impl(0)
        print(f"{M:6d} {N:5d} {Kd:5d} : {t[0]:8.1f} | {t[1]:8.1f} | {2*M*N*Kd/t[0]/1e6:7.1f}")
    convs = [(121, 213, 64, 64, 3, 1, 1), (61, 107, 128, 128, 3, 1, 1), (31, 54, 256, 256, 3, 1, 1),
             (121, 213, 128, 128, 3, 2, 1), (61, 107, 256, 256, 3, 2, 1), (121, 213, 256, 512, 1, 2, 0),
             (121, 213, 128, 128, 3, 1, 1)]
    print("conv H W Cin Cout k s : tc us | legacy us | TF/s tc")
    for H, W, Cin, Cout, k, s, pad in convs:
        x = torch.randn(H, W, Cin, device=dev).to(OP)
        w = torch.randn(Cout, k, k, Cin, device=dev).to(OP)
        b = torch.randn(Cout, device=dev)
        t = []
        for impl in (0, 1):
            lib.rmem_set_gemm_impl(impl)
            t.append(timeit(lambda: K.conv2d_nhwc(x, w, b, s, pad, act=K.ACT_RELU)))
        lib.rmem_set_gemm_impl(0)
        Ho, Wo = (H + 2 * pad - k) // s + 1, (W + 2 * pad - k) // s + 1
        print(f"{H:4d} {W:4d} {Cin:5d} {Cout:5d} {k} {s} : {t[0]:8.1f} | {t[1]:8.1f} | {2*Ho*Wo*Cout*k*k*Cin/t[0]/1e6:7.1f}")


if __name__ == "__main__":
    main()
