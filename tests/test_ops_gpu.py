"""Op-level parity: every CUDA kernel on the path, called through the C ABI, against the CPU oracle
(oracle/rmem_oracle.py, itself pinned to the reference) on the same seeded inputs.

Tolerances: tensor-core operands are 16-bit (fp16 by default, 2^-11 rounding; bf16 build 2^-9) with fp32
accumulation, the oracle is fp32 fed the same operand-rounded inputs.  rel-Frobenius <= 6e-3 per op for GEMM-like
ops (covers the bf16 build too), integer / index outputs bit-exact.
"""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import rmem_oracle as O

pytestmark = pytest.mark.gpu

from rmem_b200 import _capi  # noqa: E402

OP = _capi.op_dtype()        # 16-bit tensor-core operand type of the build (fp16 default)


def relfro(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-12))


def bfr(t):  # round to the kernels' 16-bit operand type
    return t.to(OP).float()


@pytest.fixture(scope="module")
def ops(cuda_device):
    from rmem_b200 import ops as K
    return K


def test_gemm_linear_epilogues(ops, cuda_device):
    g = torch.Generator().manual_seed(0)
    for (M, N, K) in [(1674, 640, 256), (100, 64, 32), (1674, 225, 128), (300, 1024, 1032)]:
        A = torch.randn(M, K, generator=g)
        W = torch.randn(N, K, generator=g) / math.sqrt(K)
        b = torch.randn(N, generator=g)
        ref = F.linear(bfr(A), bfr(W), b)
        out = ops.gemm(A.to(cuda_device).to(OP), W.to(cuda_device).to(OP), b.to(cuda_device), out_f32=True)
        assert relfro(out, ref) < 2e-5 * math.sqrt(K), (M, N, K)
        # silu from a column + bf16 output + gate + residual
        gate = torch.randn(M, N, generator=g)
        res = torch.randn(M, N, generator=g)
        ref2 = F.linear(bfr(A), bfr(W), b) + bfr(res)
        ref2 = torch.cat([ref2[:, : N // 2 // 2 * 2], O.silu(ref2[:, N // 2 // 2 * 2:])], 1) * bfr(gate)
        out2 = ops.gemm(A.to(cuda_device).to(OP), W.to(cuda_device).to(OP), b.to(cuda_device),
                        act=ops.ACT_SILU, act_from=N // 2 // 2 * 2, residual=res.to(cuda_device).to(OP),
                        gate=gate.to(cuda_device).to(OP))
        assert relfro(out2, ref2) < 6e-3, (M, N, K)


def test_gemm_bias_along_m_and_accumulate(ops, cuda_device):
    g = torch.Generator().manual_seed(1)
    M, N, K = 512, 1674, 256
    A = torch.randn(M, K, generator=g) / 16
    Bm = torch.randn(N, K, generator=g)
    b = torch.randn(M, generator=g)
    ref = O.silu(bfr(A) @ bfr(Bm).t() + b[:, None])
    out = ops.gemm(A.to(cuda_device).to(OP), Bm.to(cuda_device).to(OP), b.to(cuda_device), act=ops.ACT_SILU,
                   bias_along_m=True, out_f32=True)
    assert relfro(out, ref) < 1e-3
    acc = torch.randn(M, N, generator=g)
    acc_d = acc.to(cuda_device).clone()
    ops.gemm(A.to(cuda_device).to(OP), Bm.to(cuda_device).to(OP), None, accumulate_into=acc_d)
    assert relfro(acc_d, acc + bfr(A) @ bfr(Bm).t()) < 1e-4


@pytest.mark.parametrize("cfg", [(33, 41, 8, 64, 7, 2, 3), (31, 54, 256, 256, 3, 1, 1), (61, 107, 64, 128, 3, 2, 1),
                                 (61, 107, 128, 64, 1, 2, 0), (17, 19, 1024, 256, 1, 1, 0)])
def test_conv_implicit_gemm(ops, cuda_device, cfg):
    Hin, Win, Cin, Cout, k, s, p = cfg
    g = torch.Generator().manual_seed(2)
    x = torch.randn(1, Cin, Hin, Win, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(Cin * k * k)
    b = torch.randn(Cout, generator=g)
    ref = F.relu(F.conv2d(bfr(x), bfr(w), b, stride=s, padding=p))[0].permute(1, 2, 0)
    out = ops.conv2d_nhwc(x[0].permute(1, 2, 0).contiguous().to(cuda_device).to(OP),
                          w.permute(0, 2, 3, 1).contiguous().to(cuda_device).to(OP), b.to(cuda_device), s, p,
                          act=ops.ACT_RELU)
    assert tuple(out.shape) == tuple(ref.shape)
    assert relfro(out, ref) < 6e-3


def test_layernorm_groupnorm(ops, cuda_device):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(1674, 256, generator=g) * 2 + 0.3
    gm, bt = torch.rand(256, generator=g) + 0.5, torch.randn(256, generator=g)
    ref = F.layer_norm(x, (256,), gm, bt, 1e-5)
    out = ops.layernorm(x.to(cuda_device), gm.to(cuda_device), bt.to(cuda_device))
    assert relfro(out, ref) < 4e-3
    x = torch.randn(1674, 512, generator=g) * 1.5 + 0.2
    gm, bt = torch.rand(512, generator=g) + 0.5, torch.randn(512, generator=g)
    ref = O.group_norm_tokens(x, gm, bt, 2)
    out = ops.groupnorm(x.to(cuda_device), gm.to(cuda_device), bt.to(cuda_device), 2, False)
    assert relfro(out, ref) < 4e-3
    x = bfr(torch.randn(61 * 107, 128, generator=g) + 0.1)
    gm, bt = torch.rand(128, generator=g) + 0.5, torch.randn(128, generator=g)
    ref = F.relu(O.group_norm_tokens(x, gm, bt, 8))
    out = ops.groupnorm(x.to(cuda_device).to(OP), gm.to(cuda_device), bt.to(cuda_device), 8, True)
    assert relfro(out, ref) < 4e-3


def test_dwconv_upsample_maxpool_transpose(ops, cuda_device):
    g = torch.Generator().manual_seed(4)
    for (h, w, C) in [(31, 54, 1024), (17, 21, 1024), (46, 81, 1024), (5, 3, 64), (8, 18, 128)]:   # ragged tiles too
        x = bfr(torch.randn(h * w, C, generator=g))
        wt = torch.randn(C, 1, 5, 5, generator=g) / 5
        ref = O.dwconv5(x, wt, h, w)
        out = ops.dwconv5x5(x.to(cuda_device).to(OP), wt.view(C, 25).t().contiguous().to(cuda_device), h, w)
        assert relfro(out, ref) < 4e-3, (h, w, C)
    h, w = 31, 54
    xm = bfr(torch.randn(1, 256, h, w, generator=g))
    ref = F.interpolate(xm, size=(61, 107), mode="bilinear", align_corners=True)[0].permute(1, 2, 0)
    out = ops.upsample_bilinear(xm[0].permute(1, 2, 0).contiguous().to(cuda_device).to(OP), 61, 107)
    assert relfro(out, ref) < 4e-3
    xp = bfr(torch.randn(1, 64, 65, 81, generator=g))
    ref = F.max_pool2d(xp, 3, 2, 1)[0].permute(1, 2, 0)
    out = ops.maxpool3x3s2(xp[0].permute(1, 2, 0).contiguous().to(cuda_device).to(OP))
    assert torch.equal(out.float().cpu(), ref)
    xt = bfr(torch.randn(1674, 1024, generator=g))
    out = ops.transpose(xt.to(cuda_device).to(OP), 1792)
    assert torch.equal(out[:, :1674].float().cpu(), xt.t())
    assert float(out[:, 1674:].abs().max()) == 0.0


def test_id_embedding_matches_conv_of_one_hot(ops, cuda_device):
    sd = O.make_state_dict("r50_deaotl", seed=0)
    cfg = O.OracleConfig()
    H, W = 257, 321
    lab = O.synthetic_label(H, W, 10)
    lab[0, 0, 5:30, 7:50] = 255
    from rmem_b200.weights import pack_deaot
    pk = pack_deaot(sd)
    # a second label map with salt-and-pepper noise: mixed patches with a dominant class (prefix + minority correction),
    # patches without one (plain tap loop), ids above 10 that fall in no one-hot channel, and ignore pixels
    g = torch.Generator().manual_seed(9)
    noisy = lab.clone()
    m = torch.rand(lab.shape, generator=g) < 0.12
    noisy[m] = torch.randint(0, 14, lab.shape, generator=g)[m].to(noisy.dtype)
    noisy[0, 0, 100:140, 100:160] = torch.randint(0, 11, (40, 60), generator=g).to(noisy.dtype)
    noisy[0, 0, 200:215, 10:40] = 255
    for which, lb in (("blocks", lab), ("noisy", noisy)):
        for use_ignore in (False, True):
            ref = O.id_embedding(sd, cfg, O.one_hot_with_ignore(lb, use_ignore))
            # tap loop only / rectangle + dominant-class shortcuts / + row runs
            for prefix, rows in ((None, None), (pk["idbank.prefix"].to(cuda_device), None),
                                 (pk["idbank.prefix"].to(cuda_device), pk["idbank.prefix_rows"].to(cuda_device))):
                out = ops.id_embedding(lb[0, 0].to(torch.uint8).to(cuda_device), pk["idbank.w"].to(cuda_device),
                                       pk["idbank.b"].to(cuda_device), pk["id_norm.g"].to(cuda_device),
                                       pk["id_norm.b"].to(cuda_device), use_ignore, prefix=prefix, prefix_rows=rows)
                assert relfro(out, ref) < 2e-4, (which, use_ignore, prefix is not None, rows is not None)


def _attn_inputs(T, HW, g, sharp=1.0, Dv=1024):
    q = torch.randn(HW, 128, generator=g) * sharp
    k = torch.randn(T, HW, 128, generator=g)
    v = torch.randn(T, HW, Dv, generator=g)
    return bfr(q), bfr(k), bfr(v)


@pytest.mark.parametrize("T,HW,slots", [(1, 289, [0]), (3, 357, [2, 0, 3]), (8, 1674, [0, 5, 1, 2, 8, 3, 4, 6])])
def test_long_attention_dense(ops, cuda_device, T, HW, slots):
    g = torch.Generator().manual_seed(5)
    q, k, v = _attn_inputs(T, HW, g, sharp=2.0)
    pe_cur = torch.randn(128, generator=g) * 0.1
    pe_mem = torch.randn(4, 128, generator=g) * 0.5
    gate = bfr(torch.randn(HW, 1024, generator=g))
    nslots = max(slots) + 1
    qt = bfr(q + pe_cur)
    kt = k + O.temporal_pe(pe_mem, T).view(T, 1, -1)
    ref, ref_mass = O.long_term_attention(qt, kt, v, 128)
    kb, vtb, HWp = ops.build_bank(k.to(cuda_device), v.to(cuda_device), nslots, slots)
    out, mass = ops.long_attention(q.to(cuda_device).to(OP), kb, vtb, slots, HW, pe_cur.to(cuda_device),
                                   pe_mem.to(cuda_device), gate.to(cuda_device).to(OP))
    assert relfro(out, ref * gate) < 8e-3
    assert float((mass.cpu() - ref_mass).abs().max()) < 2e-3
    assert float((mass.sum(1).cpu() - 1).abs().max()) < 2e-3


def test_temporal_pe_slots_match_oracle(ops, cuda_device):
    for T in range(1, 13):
        ref = [lo if fr == 0.0 else None for lo, hi, fr in O.temporal_pe_slots(T)]
        assert ops.temporal_pe_slots(T) == ref, T


@pytest.mark.parametrize("impl", ["tc", "tc15", "cuda_core"])
def test_local_attention(ops, cuda_device, impl):
    """tc = tensor-core kernel with the packed 16-float window rows, tc15 = same kernel on the reference order."""
    kw = {"impl": "tc", "rel_pitch": 15} if impl == "tc15" else {"impl": impl}
    g = torch.Generator().manual_seed(6)
    for (h, w) in [(17, 21), (31, 54), (9, 70), (46, 81)]:
        HW = h * w
        q = bfr(torch.randn(HW, 128, generator=g))
        k = bfr(torch.randn(HW, 128, generator=g))
        v = bfr(torch.randn(HW, 1024, generator=g))
        rw = bfr(torch.randn(225, 128, generator=g) * 0.1)
        rb = torch.randn(225, generator=g) * 0.1
        gate = bfr(torch.randn(HW, 1024, generator=g))
        ref = O.local_attention(q, k, v, rw, rb, h, w) * gate
        out = ops.local_attention(q.to(cuda_device).to(OP), k.to(cuda_device).to(OP),
                                  v.to(cuda_device).to(OP), rw.to(cuda_device), rb.to(cuda_device), h, w,
                                  gate.to(cuda_device).to(OP), **kw)
        assert torch.isfinite(out.float()).all()
        assert relfro(out, ref) < 6e-3, (h, w, impl)


def test_mask_head_bit_exact_labels(ops, cuda_device):
    """Integer mask IDs must be bit-exact given identical 1/4-res logits.  Full-res logits: tight for one object
    group; for k > 1 the reference's logit(clamp(p)) is ill-conditioned near p -> 1 (1 ulp of p moves the logit by
    ~6e-3), so those are compared in probability space."""
    g = torch.Generator().manual_seed(7)
    for k, (h4, w4, Ho, Wo) in [(1, (65, 81, 257, 321)), (1, (121, 213, 480, 854)), (2, (49, 65, 193, 257)),
                                (3, (33, 41, 129, 161)), (1, (65, 81, 65, 81))]:
        lgs = [torch.randn(1, 11, h4, w4, generator=g) * 3 for _ in range(k)]
        up = [F.interpolate(l, size=(Ho, Wo), mode="bilinear", align_corners=True) for l in lgs]
        ref_logit = O.soft_logit_aggregation(up)
        ref_label = O.logits_to_label(ref_logit)[0, 0].to(torch.uint8)
        out, lab = ops.mask_head([l[0].contiguous().to(cuda_device) for l in lgs], Ho, Wo)
        if k == 1:
            err = float((out.cpu() - ref_logit[0]).abs().max())
            print(f"mask_head k=1 {h4}x{w4}->{Ho}x{Wo}: max |dlogit| = {err:.3e}")
            assert err < 2e-5
        else:
            err = float((torch.sigmoid(out.cpu()) - torch.sigmoid(ref_logit[0])).abs().max())
            print(f"mask_head k={k}: max |dprob| = {err:.3e}")
            assert err < 2e-6
        mism = int((lab.cpu() != ref_label).sum())
        print(f"mask_head k={k}: label mismatches = {mism}/{lab.numel()}")
        assert mism == 0, f"k={k}: {mism} label mismatches"


def test_evict_relevance(ops, cuda_device):
    g = torch.Generator().manual_seed(8)
    h, w, T = 17, 21, 5
    mass = torch.rand(h * w, T, generator=g)
    mass = mass / mass.sum(1, keepdim=True)
    lg = torch.randn(1, 11, 65, 81, generator=g) * 2
    fg = 1 - torch.softmax(F.interpolate(lg, size=(h, w), mode="bilinear", align_corners=True), 1)[0, 0].flatten()
    ref = (mass * fg.view(-1, 1)).sum(0)
    out = ops.evict_relevance(mass.to(cuda_device), lg[0].contiguous().to(cuda_device), h, w)
    assert float((out.cpu() - ref).abs().max() / ref.abs().max()) < 1e-5


def test_gpu_preprocess_matches_cv2_loader(ops, cuda_device):
    """rmem_preprocess_fwd against the reference's loader arithmetic (video_transforms.py:559-682): cv2.resize(float32,
    INTER_CUBIC) -> optional horizontal flip -> /255, ImageNet mean / std, HWC -> CHW.  Tolerance 2e-3 (normalised units)."""
    import cv2
    import numpy as np
    rng = np.random.RandomState(3)
    img = cv2.GaussianBlur(rng.randint(0, 255, (97, 161, 3)).astype(np.uint8), (0, 0), 1.5)
    for (nh, nw) in ((97, 161), (129, 209), (65, 113), (113, 97)):
        for flip in (False, True):
            for bgr in (True, False):
                ref = np.array(img, dtype=np.float32)
                if bgr:
                    ref = ref[:, :, [2, 1, 0]]
                if (nh, nw) != ref.shape[:2]:
                    ref = cv2.resize(ref, dsize=(nw, nh), interpolation=cv2.INTER_CUBIC)
                if flip:
                    ref = ref[:, ::-1]
                ref = (ref / 255. - (0.485, 0.456, 0.406)) / (0.229, 0.224, 0.225)
                ref = torch.from_numpy(np.ascontiguousarray(ref.transpose(2, 0, 1))).float()
                out = ops.preprocess(torch.from_numpy(img).to(cuda_device), nh, nw, bgr=bgr, flip=flip)
                err = float((out[0].cpu() - ref).abs().max())
                assert out.shape == (1, 3, nh, nw) and err < 2e-3, (nh, nw, flip, bgr, err)


@pytest.mark.parametrize("k", [1, 2])
def test_tta_head_matches_probability_averaging(ops, cuda_device, k):
    """rmem_tta_head_fwd against the evaluator's merge (evaluator.py:420-441) written out in torch: upsample each
    augmentation's logits (bilinear, align_corners), soft-aggregate the object groups, flip back, softmax, mean, argmax."""
    g = torch.Generator().manual_seed(17 + k)
    Ho, Wo = 97, 129
    sizes = [(25, 33), (25, 33), (33, 43), (33, 43)]
    flips = [False, True, False, True]
    logits = [[torch.randn(11, h4, w4, generator=g) * 3 for _ in range(k)] for (h4, w4) in sizes]
    probs = []
    for per, fl in zip(logits, flips):
        ups = [F.interpolate(t[None], size=(Ho, Wo), mode="bilinear", align_corners=True) for t in per]
        lg = O.soft_logit_aggregation(ups)
        if fl:
            lg = torch.flip(lg, dims=(3,))
        probs.append(torch.softmax(lg, dim=1))
    mean = torch.mean(torch.cat(probs, 0), 0)
    ref_lab = torch.argmax(mean, 0)
    prob, lab = ops.tta_head([[t.to(cuda_device) for t in per] for per in logits], flips, Ho, Wo, want_prob=True)
    assert float((prob.cpu() - mean).abs().max()) < 2e-5
    top2 = torch.topk(mean, 2, dim=0).values
    decided = (top2[0] - top2[1]) > 1e-4                         # away from near-ties the label is exact
    assert bool((lab.cpu().long() == ref_lab)[decided].all()) and float(decided.float().mean()) > 0.99
