"""CPU: the training-side slice: oracle loss vs the reference's recorded values and gradient
(tests/golden/train_small.npz, written by oracle/make_train_golden.py from the UNMODIFIED reference), the host logic of
rmem_b200.training.train_forward driven by the oracle engine, the loss schedules, and a numpy restatement of the exact
algorithm csrc/train_loss.cu runs (fp32 taps, 3-pass radix select of the k-th largest cross entropy, tie handling,
Jaccard coefficients, gather-form transpose of the upsampling with its candidate ranges) against autograd -- so that the
ALGORITHM of the CUDA loss head is checked here and only its transcription is left to tests/test_training_gpu.py."""
import json
import math
import os

import numpy as np
import pytest
import torch

from oracle import rmem_oracle as O
from oracle import train_oracle as TO
from rmem_b200 import training as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "train_small.npz")
pytestmark = pytest.mark.timeout(900)        # the emulated kernels are real threads on barriers: fail, never hang


@pytest.fixture(scope="module")
def gold():
    z = np.load(GOLD)
    return json.loads(str(z["meta"])), z


def test_schedules_match_reference_formulas():
    cfg = T.TrainConfig(total_steps=20000)
    P = 129 * 161
    assert T.top_k_pixels(0, P, cfg) == P                                   # loss.py:190-198: everything at step 0
    assert T.top_k_pixels(9000, P, cfg) == 4880                             # the golden's recorded k
    assert T.top_k_pixels(10 ** 6, P, cfg) == int(0.15 * P)
    assert T.aux_weight(0, cfg) == pytest.approx(1.0)
    assert T.aux_weight(3000, cfg) == pytest.approx(0.85, abs=1e-6)         # aot_engine.py:53-54
    assert T.aux_weight(30000, cfg) == 0.0
    assert T.aux_weight(5000, T.TrainConfig(total_steps=100000, aux_loss_ratio=0.1)) == pytest.approx(0.5, abs=1e-6)


def test_oracle_loss_and_gradient_reproduce_reference(gold):
    meta, z = gold
    lg = torch.from_numpy(z["lh_logits4"]).unsqueeze(0)
    for case in meta["loss_head"]:
        name = case["name"]
        gt = torch.from_numpy(z[f"lh_gt_{name}"])
        total, ce, jac, grad = TO.loss_head_with_grad(lg, gt, meta["n_obj"], case["k"])
        ref = z[f"lh_losses_{name}"]
        assert abs(total - ref[0]) < 2e-6 and abs(ce - ref[1]) < 2e-6 and abs(jac - ref[2]) < 4e-6, (name, total, ref)
        rg = torch.from_numpy(z[f"lh_grad_{name}"])
        assert float((grad[0] - rg).abs().max() / rg.abs().max()) < 1e-4, name
        assert float(grad[0, meta["n_obj"] + 1:].abs().max()) == 0.0         # unused identities get no gradient


def test_train_forward_on_oracle_engine_reproduces_reference(gold):
    meta, z = gold
    H, W, n_obj, F_ = meta["H"], meta["W"], meta["n_obj"], meta["n_frames"]
    sd = O.make_state_dict(meta["model"], seed=meta["seed"], sharpen=meta["sharpen"])
    frames = O.synthetic_frames(F_, H, W, seed=meta["seed"] + 1)
    masks = torch.from_numpy(z["masks"]).float().unsqueeze(1)
    eng = O.OracleEngine(sd, O.OracleConfig(model=meta["model"], former_mem_len=meta["former"],
                                            latter_mem_len=meta["latter"]), long_term_mem_gap=meta["gap"])
    cfg = T.TrainConfig(**meta["train_cfg"])
    with torch.no_grad():
        loss, pred, fl, boards = T.train_forward(eng, frames, masks, 1, [n_obj], step=meta["step"], cfg=cfg,
                                                 loss_fn=lambda lg, gt, n, k: TO.loss_head(lg, gt, n, k)[0],
                                                 mask_fn=TO.predict_mask)
    assert loss.ndim == 0 and abs(float(loss) - meta["ref_loss"]) < 2e-5
    assert len(fl) == F_ and all(x.shape == (1,) for x in fl)
    assert max(abs(float(a) - b) for a, b in zip(fl, meta["ref_frame_losses"])) < 2e-5
    got = torch.stack([m[0] for m in pred]).to(torch.uint8).numpy()
    assert got.shape == z["pred_masks"].shape and int((got != z["pred_masks"]).sum()) <= 2
    assert eng.aot_engines[0].long_memories_indexes == meta["ref_idx"]
    assert boards == {"image": {}, "scalar": {}}


def test_train_forward_batch_and_prev_pred_host_logic():
    """Frame-major batches and use_prev_pred on a recording fake engine: call order and which label feeds the memory."""
    H, W, B, F_ = 17, 33, 2, 4
    calls = []

    class Sub:
        pred_id_logits = torch.zeros(1, 11, 5, 9)

    class Fake:
        aot_engines = [Sub()]

        def restart_engine(self):
            calls.append(("restart",))

        def add_reference_frame(self, img, mask, obj_nums, frame_step):
            calls.append(("ref", float(img.flatten()[0]), float(mask.flatten()[0]), obj_nums[0], frame_step))

        def match_propogate_one_frame(self, img, output_size=None):
            calls.append(("prop", float(img.flatten()[0])))

        def update_memory(self, label):
            assert tuple(label.shape) == (1, 1, H, W)
            calls.append(("mem", float(label.flatten()[0])))

    frames = torch.arange(F_ * B).float().view(-1, 1, 1, 1).expand(-1, 3, H, W).contiguous()     # value = f * B + b
    masks = (100 + torch.arange(F_ * B)).float().view(-1, 1, 1, 1).expand(-1, 1, H, W).contiguous()
    seen_gt = []

    def loss_fn(lg, gt, n, k):
        seen_gt.append((float(gt.flatten()[0]), n, k))
        return torch.tensor(float(len(seen_gt)))

    def mask_fn(lg, h, w, n):
        return torch.full((h, w), 7, dtype=torch.uint8)

    cfg = T.TrainConfig(total_steps=100)
    for prev_pred in (False, True):
        calls.clear(); seen_gt.clear()
        loss, pred, fl, _ = T.train_forward(Fake(), frames, masks, B, [3, 5], step=0, use_prev_pred=prev_pred, cfg=cfg,
                                            loss_fn=loss_fn, mask_fn=mask_fn)
        for b in range(B):
            mem = (lambda f: 7.0) if prev_pred else (lambda f, b=b: 100.0 + f * B + b)
            assert calls[b * 7:(b + 1) * 7] == [
                ("restart",), ("ref", float(b), 100.0 + b, [3, 5][b], 0), ("prop", float(B + b)),
                ("mem", mem(1)), ("prop", float(2 * B + b)), ("mem", mem(2)), ("prop", float(3 * B + b))]
        assert len(calls) == 7 * B
        assert [g for g, _, _ in seen_gt] == [100.0 + f * B + b for b in range(B) for f in range(F_)]
        assert all(k == H * W for _, _, k in seen_gt) and [n for _, n, _ in seen_gt] == [3] * F_ + [5] * F_
        # loss values were 1..8 in call order (sample-major); aux = mean of frame 0, pred = mean over the other frames
        per = torch.arange(1, 9).float().view(B, F_)
        assert float(loss) == pytest.approx(float(per[:, 0].mean() + per[:, 1:].mean()))
        assert [tuple(x.shape) for x in fl] == [(B,)] * F_ and [tuple(m.shape) for m in pred] == [(B, H, W)] * F_
        assert pred[0].dtype == torch.long


# ---------------------------------------------------------------------------------------------------------------------
# numpy restatement of csrc/train_loss.cu, kernel by kernel
# ---------------------------------------------------------------------------------------------------------------------
f32 = np.float32


def taps(in_size, out_size):
    """tl `taps()`: (i0, i1, l0, l1) per destination index, fp32 arithmetic."""
    scale = f32(in_size - 1) / f32(out_size - 1) if out_size > 1 else f32(0)
    dst = np.arange(out_size, dtype=f32)
    real = (scale * dst).astype(f32)
    i0 = np.minimum(real.astype(np.int32), in_size - 1)
    lam = np.clip((real - i0.astype(f32)).astype(f32), 0, 1).astype(f32)
    i1 = i0 + (i0 < in_size - 1)
    return scale, i0, i1, (f32(1) - lam).astype(f32), lam


def radix_select(ce_bits, k):
    """tl_hist_kernel + tl_scan_kernel x 3: (threshold pattern, ties needed, ties present)."""
    prefix, k_rem, n_ties = 0, k, 0
    for ps in range(3):
        if ps == 0:
            digits = ce_bits >> 21
        elif ps == 1:
            digits = ((ce_bits >> 10) & 2047)[(ce_bits >> 21) == prefix]
        else:
            digits = (ce_bits & 1023)[(ce_bits >> 10) == prefix]
        hist = np.bincount(digits, minlength=2048)
        per = 8
        tot = hist.reshape(256, per).sum(1)
        cum, t = 0, 255
        while t > 0 and cum + tot[t] < k_rem:
            cum += tot[t]; t -= 1
        b = t * per + per - 1
        while b > t * per and cum + hist[b] < k_rem:
            cum += hist[b]; b -= 1
        prefix = b if ps == 0 else (prefix << (11 if ps == 1 else 10)) | b
        k_rem -= cum
        n_ties = int(hist[b])
    return int(prefix), int(k_rem), n_ties


def emulate_loss_head(logits4, gt, obj_num, k, grad_scale=1.0):
    C_all, h4, w4 = logits4.shape
    H, W = gt.shape
    n_ch = obj_num + 1
    sy, y0, y1, wy0, wy1 = taps(h4, H)
    sx, x0, x1, wx0, wx1 = taps(w4, W)
    L = logits4[:n_ch].astype(f32)
    top = (wx0 * L[:, y0][:, :, x0]).astype(f32) + (wx1 * L[:, y0][:, :, x1]).astype(f32)
    bot = (wx0 * L[:, y1][:, :, x0]).astype(f32) + (wx1 * L[:, y1][:, :, x1]).astype(f32)
    x = ((wy0[:, None] * top).astype(f32) + (wy1[:, None] * bot).astype(f32)).reshape(n_ch, -1)      # [C, P]
    g = gt.reshape(-1).astype(np.int64)
    m = x.max(0)
    e = np.exp(x - m).astype(f32)
    s = e.sum(0).astype(f32)
    p = (e * (f32(1) / s)).astype(f32)
    in_range = g < n_ch
    xg = np.where(in_range, x[np.minimum(g, n_ch - 1), np.arange(g.size)], 0)
    ce = np.where(in_range, np.maximum(np.log(s).astype(f32) - (xg - m), 0), 0).astype(f32)
    valid = g != 255
    I = np.array([(p[c] * (valid & (g == c))).sum(dtype=np.float64) for c in range(n_ch)])
    S = np.array([(p[c] * valid).sum(dtype=np.float64) for c in range(n_ch)])
    N = np.array([float((valid & (g == c)).sum()) for c in range(n_ch)])
    bits = ce.view(np.uint32).astype(np.int64)
    thr, k_rem, n_ties = radix_select(bits, k)
    thr_val = np.array([thr], np.uint32).view(f32)[0]
    ce_loss = (ce[bits > thr].sum(dtype=np.float64) + k_rem * float(thr_val)) / k
    present = int((N > 0).sum())
    D = S + N - I + 1e-6
    jac = float(np.where(N > 0, 1 - I / D, 0).sum() / present) if present else 0.0
    a = np.where(N > 0, -1.0 / (D * max(present, 1)), 0).astype(f32)
    b = np.where(N > 0, I / (D * D * max(present, 1)), 0).astype(f32)
    losses = (0.5 * ce_loss + 0.5 * jac, ce_loss, jac)
    # tl_grad_pixel_kernel
    w_ce = np.where(valid & in_range, np.where(bits > thr, 1.0, np.where(bits == thr, k_rem / n_ties, 0.0)) / k, 0).astype(f32)
    onehot = (np.arange(n_ch)[:, None] == g[None, :])
    q = np.where(valid[None, :], np.where(onehot, a[:, None], b[:, None]), 0).astype(f32)
    dot = (p * q).sum(0).astype(f32)
    gup = (f32(grad_scale) * f32(0.5) * (w_ce * (p - onehot) + p * (q - dot))).astype(f32).reshape(n_ch, H, W)
    # tl_grad_gather_kernel, with its candidate ranges checked against the exact footprints
    def axis_weights(in_size, out_size, scale, i0, i1, l0, l1):
        Wm = np.zeros((in_size, out_size), f32)
        Wm[i0, np.arange(out_size)] += l0
        Wm[i1, np.arange(out_size)] += l1
        for c in range(in_size):
            if scale > 0:
                lo = max(0, int(math.floor(f32(c - 1) / scale)) - 1)
                hi = min(out_size - 1, int(math.ceil(f32(c + 1) / scale)) + 1)
            else:
                lo, hi = 0, out_size - 1
            nz = np.nonzero(Wm[c])[0]
            assert nz.size == 0 or (lo <= nz.min() and nz.max() <= hi), (c, lo, hi, nz.min(), nz.max())
        return Wm
    Wy = axis_weights(h4, H, sy, y0, y1, wy0, wy1)
    Wx = axis_weights(w4, W, sx, x0, x1, wx0, wx1)
    grad = np.zeros((C_all, h4, w4), f32)
    grad[:n_ch] = np.einsum("yh,chw,xw->cyx", Wy, gup, Wx)
    return losses, grad, (thr, k_rem, n_ties)


@pytest.mark.parametrize("case", ["golden_ignore", "golden_absent", "random_small_k", "ties_at_zero", "one_object", "same_size"])
def test_cuda_algorithm_restatement_matches_autograd(gold, case):
    meta, z = gold
    g = torch.Generator().manual_seed(11)
    if case.startswith("golden"):
        name = "ignore_shrunk" if case == "golden_ignore" else "absent_all"
        lg = z["lh_logits4"].copy()
        gt = z[f"lh_gt_{name}"]
        n_obj = meta["n_obj"]
        k = [c for c in meta["loss_head"] if c["name"] == name][0]["k"]
    elif case == "random_small_k":
        lg = (3 * torch.randn(11, 19, 23, generator=g)).numpy()
        gt = torch.randint(0, 8, (73, 89), generator=g).to(torch.uint8).numpy()
        gt[5:20, 30:60] = 255
        n_obj, k = 7, 100
    elif case == "ties_at_zero":                       # most pixels ignored: the k-th largest value is the 0 of an ignored pixel
        lg = (3 * torch.randn(11, 9, 11, generator=g)).numpy()
        gt = np.full((33, 41), 255, np.uint8)
        gt[3:9, 4:30] = 1
        gt[20:22, 5:9] = 0
        n_obj, k = 2, 600
    elif case == "one_object":
        lg = torch.randn(11, 17, 17, generator=g).numpy()
        gt = (torch.rand(65, 65, generator=g) > 0.7).to(torch.uint8).numpy()
        n_obj, k = 1, 65 * 65
    else:                                               # label map at the logits' own size: scale 1, taps collapse
        lg = torch.randn(11, 21, 25, generator=g).numpy()
        gt = torch.randint(0, 4, (21, 25), generator=g).to(torch.uint8).numpy()
        n_obj, k = 3, 200
    (total, ce, jac), grad, (thr, k_rem, n_ties) = emulate_loss_head(lg.astype(np.float32), gt, n_obj, k)
    o_total, o_ce, o_jac, o_grad = TO.loss_head_with_grad(torch.from_numpy(lg).float().unsqueeze(0),
                                                          torch.from_numpy(gt), n_obj, k)
    assert 1 <= k_rem <= n_ties
    assert abs(ce - o_ce) < 5e-6 * max(1.0, o_ce) and abs(jac - o_jac) < 5e-6 and abs(total - o_total) < 5e-6 * max(1.0, o_total)
    og = o_grad[0].numpy()
    assert np.abs(grad - og).max() <= 2e-5 * np.abs(og).max() + 1e-10, (case, np.abs(grad - og).max(), np.abs(og).max())
    if case == "ties_at_zero":
        assert thr == 0 and n_ties > k_rem             # the threshold is an ignored pixel's zero: those carry no gradient


# ---------------------------------------------------------------------------------------------------------------------
# the ACTUAL kernels of csrc/train_loss.cu, executed on the CPU by the host emulation of tests/cuda_emu (real threads per
# block, barriers for __syncthreads and the warp primitives): the transcription of the algorithm above is checked here
# too, before the code ever sees a GPU.  "Device" pointers are numpy arrays.
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def emu_lib():
    import importlib.util
    spec = importlib.util.spec_from_file_location("cuda_emu_build", os.path.join(ROOT, "tests", "cuda_emu", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build("train_loss.cu")


def emu_loss_head(lib, lg, gt, n_obj, k, want_grad=True, grad_scale=1.0):
    import ctypes as C
    lg = np.ascontiguousarray(lg, np.float32)
    gt = np.ascontiguousarray(gt, np.uint8)
    Cn, h4, w4 = lg.shape
    H, W = gt.shape
    n = C.c_size_t(0)
    assert lib.rmem_train_loss_workspace_bytes(H, W, C.byref(n)) == 0
    raw = np.full(n.value + 256, 0xCD, np.uint8)                   # poisoned: the kernels must not read what they did not write
    off = (-raw.ctypes.data) % 256
    ws = raw[off:off + n.value]
    losses = np.full(3, np.nan, np.float32)
    grad = np.full_like(lg, np.nan) if want_grad else None
    vp = lambda a: C.c_void_p(None if a is None else a.ctypes.data)   # noqa: E731
    rc = lib.rmem_train_loss_fwd_bwd(vp(lg), Cn, h4, w4, vp(gt), H, W, int(n_obj), C.c_longlong(int(k)),
                                     C.c_float(grad_scale), vp(losses), vp(grad), vp(ws), C.c_size_t(n.value), None)
    return rc, losses, grad


# (the two golden label maps run in the next test, once each: the emulation spends ~1 s per call at 129 x 161)
EMU_CASES = ["random_small_k", "ties_at_zero", "one_object", "same_size", "no_objects"]


def emu_case(gold, case):
    meta, z = gold
    g = torch.Generator().manual_seed(11)
    if case.startswith("golden"):
        name = "ignore_shrunk" if case == "golden_ignore" else "absent_all"
        k = [c for c in meta["loss_head"] if c["name"] == name][0]["k"]
        return z["lh_logits4"].copy(), z[f"lh_gt_{name}"], meta["n_obj"], k
    if case == "random_small_k":
        lg = (3 * torch.randn(11, 19, 23, generator=g)).numpy()
        gt = torch.randint(0, 8, (73, 89), generator=g).to(torch.uint8).numpy()
        gt[5:20, 30:60] = 255
        return lg, gt, 7, 100
    if case == "ties_at_zero":
        lg = (3 * torch.randn(11, 9, 11, generator=g)).numpy()
        gt = np.full((33, 41), 255, np.uint8)
        gt[3:9, 4:30] = 1
        gt[20:22, 5:9] = 0
        return lg, gt, 2, 600
    if case == "one_object":
        lg = torch.randn(11, 17, 17, generator=g).numpy()
        return lg, (torch.rand(65, 65, generator=g) > 0.7).to(torch.uint8).numpy(), 1, 65 * 65
    if case == "same_size":
        lg = torch.randn(11, 21, 25, generator=g).numpy()
        return lg, torch.randint(0, 4, (21, 25), generator=g).to(torch.uint8).numpy(), 3, 200
    lg = torch.randn(11, 9, 9, generator=g).numpy()                # no_objects: one channel, everything is exactly zero
    return lg, np.zeros((33, 33), np.uint8), 0, 33 * 33


@pytest.mark.parametrize("case", EMU_CASES)
def test_cuda_kernels_on_host_emulation_vs_oracle(gold, emu_lib, case):
    lg, gt, n_obj, k = emu_case(gold, case)
    rc, losses, grad = emu_loss_head(emu_lib, lg, gt, n_obj, k)
    assert rc == 0, emu_lib.rmem_last_error()
    total, ce, jac, og = TO.loss_head_with_grad(torch.from_numpy(lg).float().unsqueeze(0), torch.from_numpy(gt), n_obj, k)
    for got, want in zip(losses.astype(np.float64), (total, ce, jac)):
        assert abs(got - want) <= 2e-5 * max(1.0, abs(want)), (case, losses, (total, ce, jac))
    og = og[0].numpy()
    assert np.isfinite(grad).all()
    assert np.abs(grad - og).max() <= 1e-4 * np.abs(og).max() + 1e-10, (case, np.abs(grad - og).max(), np.abs(og).max())
    assert not grad[n_obj + 1:].any()
    # the numpy restatement above is the same algorithm: threshold handling included, they agree far below the tolerance
    (t2, c2, j2), g2, _ = emulate_loss_head(lg.astype(np.float32), gt, n_obj, k)
    assert abs(losses[1] - c2) <= 2e-6 * max(1.0, c2) and np.abs(grad - g2).max() <= 2e-6 * np.abs(g2).max() + 1e-12
    if case not in ("ties_at_zero", "no_objects"):
        return
    # run to run bit-identical; forward-only leaves the losses unchanged; grad_scale is exact for a power of two
    rc, losses2, grad2 = emu_loss_head(emu_lib, lg, gt, n_obj, k)
    assert rc == 0 and np.array_equal(losses2, losses) and np.array_equal(grad2, grad)
    rc, losses3, none = emu_loss_head(emu_lib, lg, gt, n_obj, k, want_grad=False)
    assert rc == 0 and none is None and np.array_equal(losses3, losses)
    rc, _, grad4 = emu_loss_head(emu_lib, lg, gt, n_obj, k, grad_scale=0.25)
    assert rc == 0 and np.array_equal(grad4 * 4, grad)


def test_cuda_kernels_on_host_emulation_reference_golden_and_mask(gold, emu_lib):
    import ctypes as C
    meta, z = gold
    for case in meta["loss_head"]:
        name = case["name"]
        rc, losses, grad = emu_loss_head(emu_lib, z["lh_logits4"], z[f"lh_gt_{name}"], meta["n_obj"], case["k"])
        assert rc == 0
        ref = z[f"lh_losses_{name}"]
        assert np.abs(losses - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())
        rg = z[f"lh_grad_{name}"]
        assert np.abs(grad - rg).max() <= 1e-4 * np.abs(rg).max()
    # rmem_train_predict_mask against the reference's recorded masks would need the engine; against the oracle's argmax:
    g = torch.Generator().manual_seed(5)
    for (h4, w4, H, W, n_obj) in [(19, 23, 73, 89, 7), (21, 25, 21, 25, 3), (9, 9, 33, 33, 1)]:
        lg = (3 * torch.randn(11, h4, w4, generator=g)).numpy()
        lab = np.full((H, W), 99, np.uint8)
        rc = emu_lib.rmem_train_predict_mask(C.c_void_p(lg.ctypes.data), 11, h4, w4, H, W, n_obj,
                                             C.c_void_p(lab.ctypes.data), None)
        assert rc == 0
        want = TO.predict_mask(torch.from_numpy(lg), H, W, n_obj).numpy()
        assert int((lab != want).sum()) <= 1 and lab.max() <= n_obj


def test_cuda_entry_points_reject_bad_arguments_on_host_emulation(emu_lib):
    lg, gt = np.zeros((11, 9, 9), np.float32), np.zeros((33, 33), np.uint8)
    for n_obj, k in ((11, 10), (3, 0), (3, 33 * 33 + 1)):
        rc, _, _ = emu_loss_head(emu_lib, lg, gt, n_obj, k)
        assert rc == -1 and emu_lib.rmem_last_error()
    rc, _, _ = emu_loss_head(emu_lib, lg[:3], gt, 5, 10)
    assert rc == -1


def test_train_forward_logit_grads_are_the_gradient_of_the_returned_loss():
    """logit_grads: F tensors [B, C, h4, w4] = d loss / d pred_id_logits per frame, with the frame weights of
    aot_engine.py:104-109 (w_aux / B for the reference frame, 1 / ((F - 1) B) for the others).  Checked against autograd
    through the whole aggregation on a fake engine that serves fixed logits per (sample, frame)."""
    H, W, B, F_, n_obj = 33, 41, 2, 3, 3
    g = torch.Generator().manual_seed(2)
    logits = [[(2 * torch.randn(1, 11, 9, 11, generator=g)).requires_grad_(True) for _ in range(F_)] for _ in range(B)]
    gts = torch.randint(0, n_obj + 1, (F_ * B, 1, H, W), generator=g).float()
    gts[:, :, 3:9, 5:15] = 255
    frames = torch.zeros(F_ * B, 3, H, W)

    class Sub:
        pred_id_logits = None

    class Fake:
        def __init__(self):
            self.aot_engines = [Sub()]
            self.b, self.f = -1, 0

        def restart_engine(self):
            self.b += 1
            self.f = 0

        def add_reference_frame(self, img, mask, obj_nums, frame_step):
            self.aot_engines[0].pred_id_logits = logits[self.b][0]

        def match_propogate_one_frame(self, img, output_size=None):
            self.f += 1
            self.aot_engines[0].pred_id_logits = logits[self.b][self.f]

        def update_memory(self, label):
            pass

    cfg = T.TrainConfig(total_steps=1000)
    step = 300
    # autograd through train_forward itself (differentiable oracle loss)
    loss, _, fl, _ = T.train_forward(Fake(), frames, gts, B, [n_obj] * B, step=step, cfg=cfg,
                                     loss_fn=lambda lg, gt, n, k: TO.loss_head(lg, gt, n, k)[0], mask_fn=TO.predict_mask)
    loss.backward()
    want = [torch.stack([logits[b][f].grad[0] for b in range(B)]) for f in range(F_)]

    def loss_grad_fn(lg, gt, n, k, scale):
        total, _, _, grad = TO.loss_head_with_grad(lg.detach(), gt, n, k)
        return torch.tensor(total), scale * grad

    got = []
    loss2, _, fl2, _ = T.train_forward(Fake(), frames, gts, B, [n_obj] * B, step=step, cfg=cfg, mask_fn=TO.predict_mask,
                                       logit_grads=got, loss_grad_fn=loss_grad_fn)
    assert abs(float(loss2) - float(loss.detach())) < 1e-6 and len(got) == F_
    for f in range(F_):
        assert got[f].shape == (B, 11, 9, 11)
        assert float((got[f] - want[f]).abs().max()) <= 1e-5 * float(want[f].abs().max()) + 1e-12
    assert T.aux_weight(step, cfg) == pytest.approx(0.7, abs=1e-6)
