"""CPU: the integer / exact-arithmetic kernels of rmem_b200/csrc/ops.cu -- the mask head (bit-exact label
target), the per-engine label split (a1), the ID bank with its four decompositions (a5) and the evict relevance (a12) --
compiled UNMODIFIED for the host emulation of tests/cuda_emu and run against the same torch / oracle references the GPU
tests use (tests/test_ops_gpu.py).  The GPU tests stay the parity tests proper; these make the same kernels checkable in
the build container, which has no GPU.  Sizes are small: an emulated block is 128-256 real threads."""
import ctypes as C
import importlib.util
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import rmem_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.timeout(900)        # an emulated block is real threads on barriers: fail, never hang

SHIMS = """
extern "C" {
int emu_mask_head(const float* const* logits4, int k, int h4, int w4, int Ho, int Wo, float* out_logits, uint8_t* out_label) {
  return rmem::mask_head(logits4, k, h4, w4, Ho, Wo, out_logits, out_label, nullptr);
}
int emu_separate_label(const void* label, int is_f32, uint8_t* out, int H, int W, int engine, int n_engines) {
  return rmem::separate_label(label, is_f32, out, H, W, engine, n_engines, nullptr);
}
int emu_idbank(const uint8_t* label, int H, int W, int use_ignore, const float* w_packed, const float* bias,
               const float* ln_g, const float* ln_b, float* out_f32, int h, int w, int Cc, const float* prefix,
               const float* prefix_rows) {
  return rmem::idbank_embed(label, H, W, use_ignore, w_packed, bias, ln_g, ln_b, nullptr, 0, out_f32, h, w, Cc, nullptr,
                            prefix, prefix_rows);
}
int emu_layernorm(const float* x, const float* g, const float* b, void* y, int P, int Cc) {
  return rmem::layernorm(x, Cc, g, b, (t16*)y, Cc, nullptr, 0, P, Cc, nullptr);
}
int emu_groupnorm(const void* x, int x_is_f32, const float* g, const float* b, void* y, int P, int Cc, int G, int relu,
                  double* stats) {
  return x_is_f32 ? rmem::groupnorm_f32((const float*)x, g, b, (t16*)y, P, Cc, G, relu, stats, nullptr)
                  : rmem::groupnorm_t16((const t16*)x, g, b, (t16*)y, P, Cc, G, relu, stats, nullptr);
}
int emu_dwconv(const void* x, const float* w, void* y, int h, int wd, int Cc) {
  return rmem::dwconv5x5((const t16*)x, w, (t16*)y, h, wd, Cc, nullptr);
}
int emu_upsample(const void* x, void* y, int hin, int win, int hout, int wout, int Cc) {
  return rmem::upsample_bilinear_t16((const t16*)x, (t16*)y, hin, win, hout, wout, Cc, nullptr);
}
int emu_maxpool(const void* x, void* y, int Hin, int Win, int Cc, int Hout, int Wout) {
  return rmem::maxpool3x3s2((const t16*)x, (t16*)y, Hin, Win, Cc, Hout, Wout, nullptr);
}
int emu_transpose(const void* x, long long ldx, void* y, long long ldy, int P, int Cc) {
  return rmem::transpose_t16((const t16*)x, ldx, (t16*)y, ldy, P, Cc, nullptr);
}
int emu_gn_scratch_doubles() { return rmem::kGnScratchDoubles; }
const char* emu_operand() { return RMEM_OPERAND_NAME; }
int emu_tta_head(const float* const* logits4, int n_aug, int k, const int* h4, const int* w4, const int* flip, int Ho, int Wo,
                 float* out_prob, uint8_t* out_label) {
  return rmem::tta_head(logits4, n_aug, k, h4, w4, flip, Ho, Wo, out_prob, out_label, nullptr);
}
int emu_preprocess(const uint8_t* img, int H, int W, int bgr, int nh, int nw, int flip, float* out) {
  return rmem::preprocess_frame(img, H, W, bgr, nh, nw, flip, out, nullptr);
}
int emu_qprep(const void* q, long long ldq, const float* pe_cur, const float* pe_mem, const int* pe_slot, int T, float scale,
              void* qt, float* qbias, int P, int Cc) {
  return rmem::qprep((const t16*)q, ldq, pe_cur, pe_mem, pe_slot, T, scale, (t16*)qt, qbias, P, Cc, nullptr);
}
int emu_evict_relevance(const float* mass, int T, const float* logits4, int h4, int w4, int h, int w, float* rel) {
  return rmem::evict_relevance(mass, T, logits4, h4, w4, h, w, rel, nullptr);
}
}
"""


@pytest.fixture(scope="module")
def lib():
    spec = importlib.util.spec_from_file_location("cuda_emu_build", os.path.join(ROOT, "tests", "cuda_emu", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build("ops.cu", extra=SHIMS)


def vp(a):
    return C.c_void_p(None if a is None else a.ctypes.data)


def npf(t):
    return np.ascontiguousarray(t.detach().cpu().numpy(), dtype=np.float32)


def test_mask_head_bit_exact_labels_on_host_emulation(lib):
    """tests/test_ops_gpu.py::test_mask_head_bit_exact_labels at emulation-friendly sizes: 0 label mismatches."""
    g = torch.Generator().manual_seed(7)
    for k, (h4, w4, Ho, Wo) in [(1, (17, 21, 65, 81)), (2, (13, 17, 49, 65)), (3, (9, 11, 33, 41)), (1, (17, 21, 17, 21)),
                                (1, (13, 17, 48, 70))]:
        lgs = [torch.randn(1, 11, h4, w4, generator=g) * 3 for _ in range(k)]
        up = [F.interpolate(l, size=(Ho, Wo), mode="bilinear", align_corners=True) for l in lgs]
        ref_logit = O.soft_logit_aggregation(up)
        ref_label = O.logits_to_label(ref_logit)[0, 0].to(torch.uint8).numpy()
        arrs = [npf(l[0]) for l in lgs]
        ptrs = (C.c_void_p * k)(*[a.ctypes.data for a in arrs])
        out = np.full((1 + 10 * k, Ho, Wo), np.nan, np.float32)
        lab = np.full((Ho, Wo), 99, np.uint8)
        assert lib.emu_mask_head(ptrs, k, h4, w4, Ho, Wo, vp(out), vp(lab)) == 0, lib.rmem_last_error()
        if k == 1:
            assert float(np.abs(out - ref_logit[0].numpy()).max()) < 2e-5
        else:
            sig = lambda a: 1 / (1 + np.exp(-a.astype(np.float64)))   # noqa: E731
            assert float(np.abs(sig(out) - sig(ref_logit[0].numpy())).max()) < 2e-6
        assert int((lab != ref_label).sum()) == 0, (k, h4, w4, Ho, Wo)
        lab2 = np.full((Ho, Wo), 99, np.uint8)                        # label-only call (no full-resolution logits)
        assert lib.emu_mask_head(ptrs, k, h4, w4, Ho, Wo, None, vp(lab2)) == 0
        assert np.array_equal(lab2, lab)


def test_separate_label_on_host_emulation(lib):
    """aot_engine.py:604-628: engine e keeps ids 10e+1 .. 10e+10 renumbered from 1; a single engine passes ids through."""
    g = torch.Generator().manual_seed(3)
    H, W = 37, 45
    lab = torch.randint(0, 31, (H, W), generator=g)
    lab[3:6, 4:9] = 255
    for is_f32, src in ((0, lab.to(torch.uint8).numpy()), (1, lab.float().numpy())):
        src = np.ascontiguousarray(src)
        for n_eng in (1, 3):
            want = O.separate_mask(lab.view(1, 1, H, W).float(), n_eng)
            for e in range(n_eng):
                out = np.full((H, W), 77, np.uint8)
                assert lib.emu_separate_label(vp(src), is_f32, vp(out), H, W, e, n_eng) == 0
                w = want[e][0, 0].numpy()
                if n_eng == 1:
                    assert np.array_equal(out, lab.to(torch.uint8).numpy())
                else:
                    assert np.array_equal(out.astype(np.int64), w.astype(np.int64)), (is_f32, n_eng, e)


def test_id_bank_decompositions_on_host_emulation(lib):
    """tests/test_ops_gpu.py::test_id_embedding_matches_conv_of_one_hot: tap loop / rectangle + dominant class / row runs
    against Conv2d(one-hot) + LayerNorm of the oracle, on a blocky and a noisy label map, with and without ignore."""
    from rmem_b200.weights import pack_deaot
    sd = O.make_state_dict("r50_deaotl", seed=0)
    cfg = O.OracleConfig()
    pk = pack_deaot(sd)
    H, W = 97, 129
    h, w = (H - 1) // 16 + 1, (W - 1) // 16 + 1
    lab = O.synthetic_label(H, W, 10)
    lab[0, 0, 5:30, 7:50] = 255
    g = torch.Generator().manual_seed(9)
    noisy = lab.clone()
    m = torch.rand(lab.shape, generator=g) < 0.12
    noisy[m] = torch.randint(0, 14, lab.shape, generator=g)[m].to(noisy.dtype)
    noisy[0, 0, 40:70, 60:100] = torch.randint(0, 11, (30, 40), generator=g).to(noisy.dtype)
    noisy[0, 0, 80:90, 10:40] = 255
    wp, b, lg_, lb_ = npf(pk["idbank.w"]), npf(pk["idbank.b"]), npf(pk["id_norm.g"]), npf(pk["id_norm.b"])
    prefix, rows = npf(pk["idbank.prefix"]), npf(pk["idbank.prefix_rows"])
    Cc = b.size
    for which, lb in (("blocks", lab), ("noisy", noisy)):
        l8 = np.ascontiguousarray(lb[0, 0].to(torch.uint8).numpy())
        for use_ignore in (False, True):
            ref = O.id_embedding(sd, cfg, O.one_hot_with_ignore(lb, use_ignore)).numpy()
            for pf, pr in ((None, None), (prefix, None), (prefix, rows)):
                out = np.full((h * w, Cc), np.nan, np.float32)
                rc = lib.emu_idbank(vp(l8), H, W, int(use_ignore), vp(wp), vp(b), vp(lg_), vp(lb_), vp(out), h, w, Cc,
                                    vp(pf), vp(pr))
                assert rc == 0, lib.rmem_last_error()
                err = float(np.linalg.norm(out - ref) / np.linalg.norm(ref))
                assert err < 2e-4, (which, use_ignore, pf is not None, pr is not None, err)


def test_evict_relevance_on_host_emulation(lib):
    """tests/test_ops_gpu.py::test_evict_relevance (aot_engine.py:355-362, transformer.py:891-906)."""
    g = torch.Generator().manual_seed(8)
    h, w, T = 17, 21, 5
    mass = torch.rand(h * w, T, generator=g)
    mass = mass / mass.sum(1, keepdim=True)
    lg = torch.randn(1, 11, 65, 81, generator=g) * 2
    fg = 1 - torch.softmax(F.interpolate(lg, size=(h, w), mode="bilinear", align_corners=True), 1)[0, 0].flatten()
    ref = (mass * fg.view(-1, 1)).sum(0).numpy()
    rel = np.full(T, np.nan, np.float32)
    mass_np, lg_np = npf(mass), npf(lg[0])                 # kept alive across the call: vp() only carries the address
    assert lib.emu_evict_relevance(vp(mass_np), T, vp(lg_np), 65, 81, h, w, vp(rel)) == 0, lib.rmem_last_error()
    assert float(np.abs(rel - ref).max() / np.abs(ref).max()) < 1e-5


def relfro(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-12))


def h16(t):
    """fp16 operand-rounded copy of a tensor as (numpy float16 array, rounded fp32 torch tensor)."""
    a = np.ascontiguousarray(t.detach().numpy().astype(np.float16))
    return a, torch.from_numpy(a.astype(np.float32))


def test_norms_on_host_emulation(lib):
    """tests/test_ops_gpu.py::test_layernorm_groupnorm at small sizes: LayerNorm, fp32- and fp16-input GroupNorm (two-level
    deterministic statistics with the last-block fold) against torch / the oracle; 16-bit outputs: rel-Frobenius 4e-3."""
    lib.emu_operand.restype = C.c_char_p
    assert lib.emu_operand() == b"fp16"
    g = torch.Generator().manual_seed(3)
    x = torch.randn(90, 256, generator=g) * 2 + 0.3
    gm, bt = torch.rand(256, generator=g) + 0.5, torch.randn(256, generator=g)
    ref = F.layer_norm(x, (256,), gm, bt, 1e-5).numpy()
    xa, ga, ba = npf(x), npf(gm), npf(bt)
    y = np.zeros((90, 256), np.float16)
    assert lib.emu_layernorm(vp(xa), vp(ga), vp(ba), vp(y), 90, 256) == 0, lib.rmem_last_error()
    assert relfro(y, ref) < 4e-3
    stats = np.zeros(lib.emu_gn_scratch_doubles(), np.float64)          # element [64] = block counter, zero before first use
    x = torch.randn(130, 512, generator=g) * 1.5 + 0.2
    gm, bt = torch.rand(512, generator=g) + 0.5, torch.randn(512, generator=g)
    ref = O.group_norm_tokens(x, gm, bt, 2).numpy()
    xa, ga, ba = npf(x), npf(gm), npf(bt)
    y = np.zeros((130, 512), np.float16)
    assert lib.emu_groupnorm(vp(xa), 1, vp(ga), vp(ba), vp(y), 130, 512, 2, 0, vp(stats)) == 0, lib.rmem_last_error()
    assert relfro(y, ref) < 4e-3
    x16, xr = h16(torch.randn(11 * 13, 128, generator=g) + 0.1)
    gm, bt = torch.rand(128, generator=g) + 0.5, torch.randn(128, generator=g)
    ref = F.relu(O.group_norm_tokens(xr, gm, bt, 8)).numpy()
    ga, ba = npf(gm), npf(bt)
    y = np.zeros((11 * 13, 128), np.float16)
    for _ in range(2):                                                   # the counter re-arms itself: a second call works
        assert lib.emu_groupnorm(vp(x16), 0, vp(ga), vp(ba), vp(y), 11 * 13, 128, 8, 1, vp(stats)) == 0, lib.rmem_last_error()
        assert relfro(y, ref) < 4e-3


def test_dwconv_upsample_maxpool_transpose_on_host_emulation(lib):
    """tests/test_ops_gpu.py::test_dwconv_upsample_maxpool_transpose at small sizes (ragged tiles included)."""
    g = torch.Generator().manual_seed(4)
    for (h, w, Cc) in [(9, 20, 64), (5, 3, 64), (8, 18, 128)]:
        x16, xr = h16(torch.randn(h * w, Cc, generator=g))
        wt = torch.randn(Cc, 1, 5, 5, generator=g) / 5
        ref = O.dwconv5(xr, wt, h, w).numpy()
        wa = npf(wt.view(Cc, 25).t().contiguous())
        y = np.zeros((h * w, Cc), np.float16)
        assert lib.emu_dwconv(vp(x16), vp(wa), vp(y), h, w, Cc) == 0, lib.rmem_last_error()
        assert relfro(y, ref) < 4e-3, (h, w, Cc)
    h, w = 7, 10
    x16, xr = h16(torch.randn(h, w, 64, generator=g))
    ref = F.interpolate(xr.permute(2, 0, 1)[None], size=(13, 19), mode="bilinear", align_corners=True)[0].permute(1, 2, 0)
    y = np.zeros((13, 19, 64), np.float16)
    assert lib.emu_upsample(vp(x16), vp(y), h, w, 13, 19, 64) == 0, lib.rmem_last_error()
    assert relfro(y, ref.numpy()) < 4e-3
    x16, xr = h16(torch.randn(17, 21, 64, generator=g))
    ref = F.max_pool2d(xr.permute(2, 0, 1)[None], 3, 2, 1)[0].permute(1, 2, 0).numpy()
    y = np.zeros((9, 11, 64), np.float16)
    assert lib.emu_maxpool(vp(x16), vp(y), 17, 21, 64, 9, 11) == 0, lib.rmem_last_error()
    assert np.array_equal(y.astype(np.float32), ref)
    x16, xr = h16(torch.randn(70, 128, generator=g))
    y = np.full((128, 96), 1, np.float16)
    assert lib.emu_transpose(vp(x16), 128, vp(y), 96, 70, 128) == 0, lib.rmem_last_error()
    assert np.array_equal(y[:, :70].astype(np.float32), xr.t().numpy())


def test_tta_head_and_preprocess_on_host_emulation(lib):
    """tests/test_ops_gpu.py::test_tta_head_matches_probability_averaging / test_gpu_preprocess_matches_cv2_loader at small
    sizes (evaluator.py:420-441; video_transforms.py:559-682)."""
    for k in (1, 2):
        g = torch.Generator().manual_seed(17 + k)
        Ho, Wo = 41, 53
        sizes = [(11, 14), (11, 14), (14, 18), (14, 18)]
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
        ref_lab = torch.argmax(mean, 0).numpy()
        arrs = [npf(t) for per in logits for t in per]
        ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        n_aug = len(sizes)
        h4 = (C.c_int * n_aug)(*[s[0] for s in sizes])
        w4 = (C.c_int * n_aug)(*[s[1] for s in sizes])
        fl = (C.c_int * n_aug)(*[int(f) for f in flips])
        prob = np.full((1 + 10 * k, Ho, Wo), np.nan, np.float32)
        lab = np.full((Ho, Wo), 99, np.uint8)
        assert lib.emu_tta_head(ptrs, n_aug, k, h4, w4, fl, Ho, Wo, vp(prob), vp(lab)) == 0, lib.rmem_last_error()
        assert float(np.abs(prob - mean.numpy()).max()) < 2e-5
        top2 = torch.topk(mean, 2, dim=0).values
        decided = ((top2[0] - top2[1]) > 1e-4).numpy()
        assert bool((lab.astype(np.int64) == ref_lab)[decided].all()) and decided.mean() > 0.99
    import cv2
    rng = np.random.RandomState(3)
    img = cv2.GaussianBlur(rng.randint(0, 255, (33, 57, 3)).astype(np.uint8), (0, 0), 1.5)
    for (nh, nw) in ((33, 57), (49, 81), (17, 33)):
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
                ref = np.ascontiguousarray(ref.transpose(2, 0, 1)).astype(np.float32)
                out = np.full((3, nh, nw), np.nan, np.float32)
                src = np.ascontiguousarray(img)
                assert lib.emu_preprocess(vp(src), 33, 57, int(bgr), nh, nw, int(flip), vp(out)) == 0, lib.rmem_last_error()
                assert float(np.abs(out - ref).max()) < 2e-3, (nh, nw, flip, bgr)


def test_temporal_pe_as_score_bias_on_host_emulation(lib):
    """a11 (transformer.py:1140-1175): K[t] += mem_pos_emb[slot(t)], Q += cur_pos_emb.  The engine never rewrites the K bank:
    qprep_kernel emits Qt = t16(Q + pe_cur) and the per-(query, frame) bias scale * <Qt_i, pe_mem[slot(t)]>, which the
    attention kernels add to the scores -- algebraically <Qt_i, K_j + pe> = <Qt_i, K_j> + <Qt_i, pe>.  Checked against the
    oracle's slot table and the explicit sums."""
    g = torch.Generator().manual_seed(6)
    P, Cc = 75, 128
    for T in (1, 4, 7, 9):
        slots = [lo for lo, hi, fr in O.temporal_pe_slots(T)]
        assert all(fr == 0.0 for _, _, fr in O.temporal_pe_slots(T))
        q16, qr = h16(torch.randn(P, Cc, generator=g))
        pe_cur = torch.randn(Cc, generator=g) * 0.1
        pe_mem = torch.randn(4, Cc, generator=g) * 0.5
        scale = 1.0 / (Cc ** 0.5)
        qt_ref = (qr + pe_cur).numpy().astype(np.float16)
        bias_ref = scale * (torch.from_numpy(qt_ref.astype(np.float32)) @ O.temporal_pe(pe_mem, T).view(T, Cc).t())
        pc, pm = npf(pe_cur), npf(pe_mem)
        sl = (C.c_int * T)(*slots)
        qt = np.zeros((P, Cc), np.float16)
        qb = np.full((P, T), np.nan, np.float32)
        assert lib.emu_qprep(vp(q16), C.c_longlong(Cc), vp(pc), vp(pm), sl, T, C.c_float(scale), vp(qt), vp(qb), P, Cc) == 0, \
            lib.rmem_last_error()
        assert np.array_equal(qt, qt_ref)
        assert float(np.abs(qb - bias_ref.numpy()).max()) < 1e-5 * max(1.0, float(bias_ref.abs().max()))
