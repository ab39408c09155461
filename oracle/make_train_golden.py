"""Pin the training-side slice against the UNMODIFIED reference; write tests/golden/train_small.npz.

TEST INFRASTRUCTURE.  Runs only in the build container (needs /root/reference, read-only):

    python oracle/make_train_golden.py

The reference's TRAINING engine (`build_engine(..., phase='train')` = DeAOTEngine, networks/engines/__init__.py:5-21) is
imported in place with the shims of oracle/make_golden.py and put in eval() mode, so that dropout / drop-path are off
and the forward is deterministic; `restart_engine(1, False)` switches the random identity shuffle off.  Then

  1. `AOTEngine.forward` (aot_engine.py:40-128) runs on a synthetic training sample (reference frame + 3 frames, moving
     rectangles with an ignore block) -> loss, per-frame losses, predicted masks.  rmem_b200.training.train_forward driven
     by the ORACLE engine and the ORACLE loss must reproduce them (asserted here, re-checked by tests/test_training.py).
  2. `AOTEngine.calculate_current_loss` (aot_engine.py:484-511) with the reference's own CrossEntropyLoss /
     SoftJaccordLoss modules is differentiated by autograd with respect to pred_id_logits on two label maps (one with an
     absent class and all pixels in the top-k, one with an ignore block and a shrunken top-k) -> loss values and the
     gradient.  oracle.train_oracle.loss_head_with_grad must reproduce them (asserted here); the CUDA loss head is then
     tested against these vectors on the GPU.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import make_golden as MG  # noqa: E402
from oracle import rmem_oracle as O  # noqa: E402
from oracle import train_oracle as TO  # noqa: E402
from rmem_b200 import training as T  # noqa: E402

CASE = dict(model="r50_deaotl", seed=7, sharpen=4.0, H=129, W=161, n_obj=3, n_frames=4, former=1, latter=2, gap=2,
            step=3000, total_steps=20000)


def training_masks(H: int, W: int, n_obj: int, n_frames: int) -> torch.Tensor:
    """[F,1,H,W] float label maps: the rectangles of synthetic_label drifting by (2, 3) pixels per frame; from frame 1
    on a 255 block (VOST-style ignore region) in the lower right corner."""
    base = O.synthetic_label(H, W, n_obj)
    out = []
    for f in range(n_frames):
        m = torch.roll(base, shifts=(2 * f, 3 * f), dims=(2, 3)).clone()
        if f >= 1:
            m[0, 0, H - 24:H - 8, W - 40:W - 10] = 255
        out.append(m)
    return torch.cat(out, 0)


def build_train_reference(sd, c):
    DefaultEngineConfig, build_vos_model, build_engine, seen = MG.import_reference()
    cfg = DefaultEngineConfig("golden", c["model"])
    cfg.MODEL_LINEAR_Q = False
    cfg.MODEL_IGNORE_TOKEN = True
    cfg.FORMER_MEM_LEN, cfg.LATTER_MEM_LEN = c["former"], c["latter"]
    cfg.TRAIN_TOTAL_STEPS = c["total_steps"]                 # configs/pre_vost.py:13
    net = build_vos_model(cfg.MODEL_VOS, cfg).eval()
    net.load_state_dict(sd, strict=True)
    seen.clear()
    eng = build_engine(cfg.MODEL_ENGINE, phase="train", aot_model=net, gpu_id=0, long_term_mem_gap=c["gap"]).eval()
    return cfg, eng


def main():
    torch.set_num_threads(8)
    c = CASE
    H, W, n_obj, F_ = c["H"], c["W"], c["n_obj"], c["n_frames"]
    sd = O.make_state_dict(c["model"], seed=c["seed"], sharpen=c["sharpen"])
    frames = O.synthetic_frames(F_, H, W, seed=c["seed"] + 1)
    masks = training_masks(H, W, n_obj, F_)
    cfg, eng = build_train_reference(sd, c)
    tcfg = T.TrainConfig(total_steps=cfg.TRAIN_TOTAL_STEPS, top_k_percent_pixels=cfg.TRAIN_TOP_K_PERCENT_PIXELS,
                         hard_mining_ratio=cfg.TRAIN_HARD_MINING_RATIO, aux_loss_weight=cfg.TRAIN_AUX_LOSS_WEIGHT,
                         aux_loss_ratio=cfg.TRAIN_AUX_LOSS_RATIO)

    # ---- 1. AOTEngine.forward of the reference ----
    sink = io.StringIO()
    with torch.no_grad(), contextlib.redirect_stdout(sink):
        eng.restart_engine(1, False)
        loss, pred_masks, frame_losses, _ = eng(frames, masks, 1, obj_nums=[n_obj], step=c["step"])
    ref_loss = float(loss)
    ref_frame_losses = [float(x) for x in frame_losses]
    ref_pred = torch.stack([m[0] for m in pred_masks]).to(torch.uint8)          # [F,H,W]
    ref_idx = list(eng.long_memories_indexes)
    print(f"reference forward: loss {ref_loss:.6f}  frame losses {[round(x, 5) for x in ref_frame_losses]}  "
          f"long_memories_indexes {ref_idx}")

    orc = O.OracleEngine(sd, O.OracleConfig(model=c["model"], former_mem_len=c["former"], latter_mem_len=c["latter"]),
                         long_term_mem_gap=c["gap"])
    with torch.no_grad():
        o_loss, o_pred, o_fl, _ = T.train_forward(
            orc, frames, masks, 1, [n_obj], step=c["step"], cfg=tcfg,
            loss_fn=lambda lg, gt, n, k: TO.loss_head(lg, gt, n, k)[0], mask_fn=TO.predict_mask)
    o_pred = torch.stack([m[0] for m in o_pred]).to(torch.uint8)
    d_loss = abs(float(o_loss) - ref_loss)
    d_fl = max(abs(float(a) - b) for a, b in zip(o_fl, ref_frame_losses))
    mism = int((o_pred != ref_pred).sum())
    print(f"train_forward(oracle engine, oracle loss) vs reference: |loss| {d_loss:.2e}  frame losses {d_fl:.2e}  "
          f"mask mismatches {mism}/{ref_pred.numel()}  idx {orc.aot_engines[0].long_memories_indexes}")
    assert d_loss < 2e-5 and d_fl < 2e-5 and mism <= 2 and orc.aot_engines[0].long_memories_indexes == ref_idx

    # ---- 2. the loss head and its gradient, by autograd through the reference's own loss modules ----
    P = H * W
    lh_cases = []
    with torch.no_grad(), contextlib.redirect_stdout(sink):
        eng.restart_engine(1, False)
        eng(frames, masks, 1, obj_nums=[n_obj], step=c["step"])
    logits4 = eng.pred_id_logits.detach().clone()                               # the last frame's [1,11,h4,w4]
    gt_a = masks[F_ - 1, 0].clone()                                             # ignore block, three objects
    gt_b = masks[0, 0].clone()
    gt_b[gt_b == 2] = 0                                                         # class 2 owns no pixel
    for name, gt, step in (("ignore_shrunk", gt_a, 9000), ("absent_all", gt_b, 0)):
        k = T.top_k_pixels(step, P, tcfg)
        x = logits4.clone().requires_grad_(True)
        eng.pred_id_logits = x
        eng.obj_nums = [n_obj]
        with contextlib.redirect_stdout(sink):
            total = eng.calculate_current_loss(gt.view(1, 1, H, W), step)
            ce = eng.losses[0]([torch.nn.functional.interpolate(x, size=(H, W), mode="bilinear", align_corners=True)
                                [0, :n_obj + 1].unsqueeze(0)], [gt.view(1, H, W).long()], step)
        total.sum().backward()
        r_total, r_ce = float(total), float(ce)
        r_jac = 2 * r_total - r_ce
        o_total, o_ce, o_jac, o_grad = TO.loss_head_with_grad(logits4, gt, n_obj, k)
        gerr = float((o_grad - x.grad).abs().max() / x.grad.abs().max())
        print(f"loss head [{name}] k={k}/{P}: reference total {r_total:.6f} ce {r_ce:.6f} jaccard {r_jac:.6f}; oracle "
              f"d_total {abs(o_total - r_total):.2e} d_ce {abs(o_ce - r_ce):.2e}; gradient rel err {gerr:.2e} "
              f"(max |g| {float(x.grad.abs().max()):.3e})")
        assert abs(o_total - r_total) < 2e-6 and abs(o_ce - r_ce) < 2e-6 and gerr < 1e-4
        lh_cases.append(dict(name=name, gt=gt.to(torch.uint8).numpy(), k=k, step=step,
                             losses=np.array([r_total, r_ce, r_jac], np.float64), grad=x.grad[0].numpy().copy()))

    meta = dict(case="train_small", **c, ref_loss=ref_loss, ref_frame_losses=ref_frame_losses, ref_idx=ref_idx,
                train_cfg=dict(total_steps=tcfg.total_steps, top_k_percent_pixels=tcfg.top_k_percent_pixels,
                               hard_mining_ratio=tcfg.hard_mining_ratio, aux_loss_weight=tcfg.aux_loss_weight,
                               aux_loss_ratio=tcfg.aux_loss_ratio),
                loss_head=[dict(name=x["name"], k=x["k"], step=x["step"]) for x in lh_cases],
                reference_commit="431cde18", torch=torch.__version__)
    arrays = dict(masks=masks[:, 0].to(torch.uint8).numpy(), pred_masks=ref_pred.numpy(),
                  lh_logits4=logits4[0].numpy())
    for x in lh_cases:
        arrays[f"lh_gt_{x['name']}"] = x["gt"]
        arrays[f"lh_losses_{x['name']}"] = x["losses"]
        arrays[f"lh_grad_{x['name']}"] = x["grad"]
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "train_small.npz"), meta=json.dumps(meta), **arrays)
    print("tests/golden/train_small.npz written")


if __name__ == "__main__":
    main()
