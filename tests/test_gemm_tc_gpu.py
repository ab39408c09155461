"""Parity of the tcgen05/TMA GEMM + implicit-GEMM conv kernel (rmem_b200/csrc/gemm_tc.cu) through the C ABI against
(a) torch fp32 on the same fp16-rounded operands and (b) the legacy mma.sync kernel, at the shapes the engine launches.
Tolerance: fp16 operands, fp32 accumulate -> rel-Frobenius <= 2e-5*sqrt(K) on fp32 outputs, <= 1e-3 on fp16 outputs."""
import ctypes as C
import math

import pytest
import torch
import torch.nn.functional as F

from rmem_b200 import _capi

pytestmark = pytest.mark.gpu


def relfro(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-12))


@pytest.fixture(scope="module")
def K(cuda_device):
    from rmem_b200 import ops
    return ops


def both(fn):
    lib = _capi.load()
    outs = []
    for impl in (0, 1):
        prev = lib.rmem_set_gemm_impl(impl)
        try:
            outs.append(fn())
        finally:
            lib.rmem_set_gemm_impl(prev)
    return outs


@pytest.mark.parametrize("shape", [(1674, 640, 256), (1674, 512, 1024), (25773, 64, 256), (25773, 256, 64),
                                   (25773, 256, 512), (3726, 1152, 512),     # wide tiles (BN = 256)
                                   (6527, 512, 128), (129, 64, 64), (1674, 128, 512), (1674, 256, 128),
                                   (512, 1674, 256),
                                   # split-K clusters (few tiles, long K): S = 4, S = 4 with 2 tiles, odd k-block count
                                   (1674, 128, 1024), (200, 64, 2304), (300, 64, 576), (1674, 256, 2304)])
def test_linear_matches_torch_and_legacy(K, cuda_device, shape):
    M, N, Kd = shape
    OP = _capi.op_dtype()
    g = torch.Generator().manual_seed(M + N + Kd)
    A = (torch.randn(M, Kd, generator=g)).to(cuda_device).to(OP)
    W = (torch.randn(N, Kd, generator=g) / math.sqrt(Kd)).to(cuda_device).to(OP)
    b = torch.randn(N, generator=g).to(cuda_device)
    res = torch.randn(M, N, generator=g).to(cuda_device).to(OP)
    ref = F.linear(A.float(), W.float(), b)
    ldc_ok = N % 8 == 0
    tc, leg = both(lambda: K.gemm(A, W, b, out_f32=True))
    assert relfro(tc, ref) < 2e-5 * math.sqrt(Kd)
    assert relfro(tc, leg) < 2e-5 * math.sqrt(Kd)
    if ldc_ok:
        ref2 = torch.relu(ref + res.float())
        tc2, leg2 = both(lambda: K.gemm(A, W, b, act=K.ACT_RELU, residual=res))
        assert relfro(tc2, ref2) < 1e-3
        assert relfro(tc2, leg2) < 1e-3


@pytest.mark.parametrize("cfg", [(31, 54, 256, 256, 3, 1, 1), (121, 213, 64, 64, 3, 1, 1), (61, 107, 128, 128, 3, 1, 1),
                                 (121, 213, 128, 128, 3, 2, 1), (61, 107, 256, 256, 3, 2, 1),
                                 (121, 213, 256, 512, 1, 2, 0), (61, 107, 512, 1024, 1, 2, 0),
                                 (31, 54, 1024, 256, 1, 1, 0), (17, 17, 64, 64, 3, 1, 1), (9, 200, 64, 64, 3, 1, 1)])
def test_conv_matches_torch_and_legacy(K, cuda_device, cfg):
    H, W, Cin, Cout, k, stride, pad = cfg
    OP = _capi.op_dtype()
    g = torch.Generator().manual_seed(H * W + Cin)
    x = torch.randn(H, W, Cin, generator=g).to(cuda_device).to(OP)
    w = (torch.randn(Cout, k, k, Cin, generator=g) / math.sqrt(k * k * Cin)).to(cuda_device).to(OP)
    b = torch.randn(Cout, generator=g).to(cuda_device)
    ref = F.conv2d(x.float().permute(2, 0, 1)[None], w.float().permute(0, 3, 1, 2), b, stride=stride, padding=pad)
    ref = torch.relu(ref)[0].permute(1, 2, 0)
    tc, leg = both(lambda: K.conv2d_nhwc(x, w, b, stride, pad, act=K.ACT_RELU))
    assert tc.shape == ref.shape
    assert relfro(tc, ref) < 1.5e-3, cfg
    assert relfro(tc, leg) < 1.5e-3, cfg


@pytest.mark.parametrize("cfg", [(31, 54, 256, 256, 3, 1, 1, 2), (61, 107, 128, 128, 3, 2, 1, 2), (121, 213, 64, 64, 3, 1, 1, 2),
                                 (121, 213, 256, 512, 1, 2, 0, 2), (17, 17, 64, 64, 3, 1, 1, 3)])
def test_stacked_image_conv_matches_per_image_conv(K, cuda_device, cfg):
    """n images stacked along M (4-D tensor map, rmem_gemm_desc.n_images; the pair encoder of rmem_engine_prefetch2): the
    halo of one image never reads its neighbour, results agree with torch per image and with the single-image launch."""
    H, W, Cin, Cout, k, stride, pad, n = cfg
    OP = _capi.op_dtype()
    g = torch.Generator().manual_seed(H * W + Cin + n)
    x = torch.randn(n, H, W, Cin, generator=g).to(cuda_device).to(OP)
    w = (torch.randn(Cout, k, k, Cin, generator=g) / math.sqrt(k * k * Cin)).to(cuda_device).to(OP)
    b = torch.randn(Cout, generator=g).to(cuda_device)
    pre = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), b, stride=stride, padding=pad)
    pre = pre.permute(0, 2, 3, 1)
    ref = torch.relu(pre)
    res = torch.randn(*ref.shape, generator=g).to(cuda_device).to(OP)
    out = K.conv2d_nhwc(x, w, b, stride, pad, act=K.ACT_RELU)
    assert out.shape == ref.shape
    assert relfro(out, ref) < 1.5e-3, cfg
    for i in range(n):
        one = K.conv2d_nhwc(x[i].contiguous(), w, b, stride, pad, act=K.ACT_RELU)
        assert relfro(out[i], one) < 1e-3, (cfg, i)
    out2 = K.conv2d_nhwc(x, w, b, stride, pad, act=K.ACT_RELU, residual=res)
    assert relfro(out2, torch.relu(pre + res.float())) < 1.5e-3, cfg


def test_splitk_is_bit_reproducible_and_matches_unsplit(K, cuda_device, monkeypatch):
    """The cluster split-K reduction adds the partial tiles in rank order: two runs give identical bits, and the
    result agrees with torch fp32 like the unsplit kernel does."""
    OP = _capi.op_dtype()
    g = torch.Generator().manual_seed(7)
    M, N, Kd = 1620, 256, 2304
    A = torch.randn(M, Kd, generator=g).to(cuda_device).to(OP)
    W = (torch.randn(N, Kd, generator=g) / math.sqrt(Kd)).to(cuda_device).to(OP)
    b = torch.randn(N, generator=g).to(cuda_device)
    o1 = K.gemm(A, W, b, out_f32=True)
    o2 = K.gemm(A, W, b, out_f32=True)
    assert torch.equal(o1, o2)
    assert relfro(o1, F.linear(A.float(), W.float(), b)) < 2e-5 * math.sqrt(Kd)


@pytest.mark.parametrize("H,W", [(257, 321), (481, 849), (129, 161)])
def test_stem_conv_matches_torch(K, cuda_device, H, W):
    """conv1 (7x7 / stride 2 / pad 3, FrozenBN folded, ReLU) as the "stem" mode of the tcgen05 GEMM (one k-block per
    window row over the zero-padded NHWC8 image) against F.conv2d in fp32; resnet.py:178-181."""
    from oracle import rmem_oracle as O
    from rmem_b200.weights import pack_deaot
    from rmem_b200 import _capi
    sd = O.make_state_dict("r50_deaotl", seed=1)
    pk = pack_deaot(sd)
    g = torch.Generator().manual_seed(5)
    img = torch.randn(1, 3, H, W, generator=g)
    OP = _capi.op_dtype()
    x16 = img.to(OP).float()                                  # what the kernel sees after packing
    w = pk["enc.conv1.w"].float()[:, :, :7, :3].permute(0, 3, 1, 2).contiguous()     # [64,3,7,7] with BN folded
    ref = F.relu(F.conv2d(x16, w, pk["enc.conv1.b"], stride=2, padding=3))[0].permute(1, 2, 0)
    out = K.stem_conv(img.to(cuda_device), pk["enc.conv1.w"].to(cuda_device), pk["enc.conv1.b"].to(cuda_device))
    err = float((out.float().cpu() - ref).abs().max() / ref.abs().max())
    assert out.shape == ref.shape and err < 4e-3, err
