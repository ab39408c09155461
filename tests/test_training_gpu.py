"""GPU: the training-side slice through the C ABI: rmem_train_loss_fwd_bwd and
rmem_train_predict_mask against the reference's recorded loss values / autograd gradient (tests/golden/train_small.npz)
and against the CPU oracle (oracle/train_oracle.py) on seeded inputs incl. the edge cases the CPU restatement covers
(ignore blocks, absent classes, k = all pixels, ties at zero, label map at the logits' size), bit reproducibility, the
c3 geometry (481 x 849, 10 objects), and rmem_b200.training.train_forward on the CUDA engine against the reference's
recorded training forward.

Tolerances: the loss head is fp32 with IEEE exp / log (no 16-bit operands): loss values within 2e-5 relative, gradient
within 1e-4 of its largest entry.  train_forward inherits the engine's fp16-operand logits (<= 1.5e-2 of the largest
logit, tests/test_engine_gpu.py): per-frame losses within 3e-3 absolute of the reference's (achieved 2e-4 on B200,
profiles/r02_train_pytest_gpu.log), masks >= 99.9 % identical (achieved 99.998 %)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import rmem_oracle as O
from oracle import train_oracle as TO

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "train_small.npz")


@pytest.fixture(scope="module")
def gold():
    z = np.load(GOLD)
    return json.loads(str(z["meta"])), z


@pytest.fixture(scope="module")
def T(cuda_device):
    from rmem_b200 import training
    return training


def check_against(head, lg, gt, n_obj, k, ref_losses, ref_grad, dev, tag=None):
    losses, grad = head(lg.to(dev), gt.to(dev), n_obj, k)
    losses, grad = losses.cpu().double().numpy(), grad.cpu()
    for got, want in zip(losses, ref_losses):
        assert abs(got - want) <= 2e-5 * max(1.0, abs(want)), (losses, ref_losses)
    scale = float(ref_grad.abs().max())
    gerr = float((grad.reshape(ref_grad.shape) - ref_grad).abs().max())
    assert gerr <= 1e-4 * scale + 1e-10
    assert torch.isfinite(grad).all()
    if tag:
        from parity_report import report
        report("training/loss_head/" + tag, losses=[float(x) for x in losses], ref_losses=[float(x) for x in ref_losses],
               grad_err_rel_to_max=gerr / scale if scale else 0.0, tolerances={"loss_rel": 2e-5, "grad_rel_to_max": 1e-4})
    return losses, grad


def test_loss_head_reproduces_reference_values_and_gradient(T, gold, cuda_device):
    meta, z = gold
    head = T.LossHead(cuda_device)
    lg = torch.from_numpy(z["lh_logits4"])
    for case in meta["loss_head"]:
        name = case["name"]
        check_against(head, lg, torch.from_numpy(z[f"lh_gt_{name}"]), meta["n_obj"], case["k"],
                      z[f"lh_losses_{name}"], torch.from_numpy(z[f"lh_grad_{name}"]), cuda_device,
                      tag="reference_" + name)


CASES = ["random_small_k", "ties_at_zero", "one_object", "same_size", "no_objects", "c3_geometry"]


def make_case(case):
    g = torch.Generator().manual_seed(11)
    if case == "random_small_k":
        lg = 3 * torch.randn(11, 19, 23, generator=g)
        gt = torch.randint(0, 8, (73, 89), generator=g).to(torch.uint8)
        gt[5:20, 30:60] = 255
        return lg, gt, 7, 100
    if case == "ties_at_zero":
        lg = 3 * torch.randn(11, 9, 11, generator=g)
        gt = torch.full((33, 41), 255, dtype=torch.uint8)
        gt[3:9, 4:30] = 1
        gt[20:22, 5:9] = 0
        return lg, gt, 2, 600
    if case == "one_object":
        lg = torch.randn(11, 17, 17, generator=g)
        return lg, (torch.rand(65, 65, generator=g) > 0.7).to(torch.uint8), 1, 65 * 65
    if case == "same_size":
        lg = torch.randn(11, 21, 25, generator=g)
        return lg, torch.randint(0, 4, (21, 25), generator=g).to(torch.uint8), 3, 200
    if case == "no_objects":                      # obj_num 0: one channel, every loss and gradient is exactly zero
        lg = torch.randn(11, 9, 9, generator=g)
        return lg, torch.zeros(33, 33, dtype=torch.uint8), 0, 33 * 33
    lg = 4 * torch.randn(11, 121, 213, generator=g)                   # c3: 481 x 849, 10 objects
    gt = O.synthetic_label(481, 849, 10)[0, 0].to(torch.uint8)
    gt[400:440, 600:800] = 255
    return lg, gt, 10, int(0.4 * 481 * 849)


@pytest.mark.parametrize("case", CASES)
def test_loss_head_vs_oracle(T, cuda_device, case):
    lg, gt, n_obj, k = make_case(case)
    total, ce, jac, grad = TO.loss_head_with_grad(lg.unsqueeze(0), gt, n_obj, k)
    head = T.LossHead(cuda_device)
    losses, g1 = check_against(head, lg, gt, n_obj, k, [total, ce, jac], grad[0], cuda_device, tag="oracle_" + case)
    # channels above obj_num carry no gradient; a second run is bit-identical (fixed-order reductions, no float atomics)
    assert float(g1[n_obj + 1:].abs().max()) == 0.0 if n_obj < 10 else True
    losses2, g2 = head(lg.to(cuda_device), gt.to(cuda_device), n_obj, k)
    assert torch.equal(g2.cpu(), g1) and np.array_equal(losses2.cpu().double().numpy(), losses)
    # forward-only call: same losses; grad_scale scales the gradient exactly (power of two)
    losses3, none = head(lg.to(cuda_device), gt.to(cuda_device), n_obj, k, want_grad=False)
    assert none is None and np.array_equal(losses3.cpu().double().numpy(), losses)
    _, g4 = head(lg.to(cuda_device), gt.to(cuda_device), n_obj, k, grad_scale=0.25)
    assert torch.equal(g4.cpu() * 4, g1)


def test_loss_head_rejects_bad_arguments(T, cuda_device):
    from rmem_b200 import _capi
    head = T.LossHead(cuda_device)
    lg, gt = torch.zeros(11, 9, 9), torch.zeros(33, 33, dtype=torch.uint8)
    for n_obj, k in ((11, 10), (3, 0), (3, 33 * 33 + 1)):
        with pytest.raises(_capi.RmemError):
            head(lg.to(cuda_device), gt.to(cuda_device), n_obj, k)
    with pytest.raises(_capi.RmemError):
        head(torch.zeros(3, 9, 9, device=cuda_device), gt.to(cuda_device), 5, 10)    # 6 channels needed, 3 given


@pytest.mark.parametrize("case", ["random_small_k", "same_size", "c3_geometry"])
def test_predict_mask_vs_oracle(T, cuda_device, case):
    lg, gt, n_obj, _ = make_case(case)
    H, W = gt.shape
    for n in sorted({n_obj, 1}):
        got = T.predict_mask(lg.to(cuda_device), H, W, n).cpu().long()
        want = TO.predict_mask(lg, H, W, n)
        assert got.max() <= n
        assert int((got != want).sum()) <= max(2, int(1e-5 * H * W)), case        # argmax near-ties only


def test_train_forward_on_cuda_engine_vs_reference(T, gold, cuda_device):
    from rmem_b200.engine import DeAOTModel, RmemConfig, build_engine
    meta, z = gold
    H, W, n_obj, F_ = meta["H"], meta["W"], meta["n_obj"], meta["n_frames"]
    sd = O.make_state_dict(meta["model"], seed=meta["seed"], sharpen=meta["sharpen"])
    frames = O.synthetic_frames(F_, H, W, seed=meta["seed"] + 1).to(cuda_device)
    masks = torch.from_numpy(z["masks"]).float().unsqueeze(1).to(cuda_device)
    cfg = RmemConfig(former_mem_len=meta["former"], latter_mem_len=meta["latter"])
    eng = build_engine("deaotengine", aot_model=DeAOTModel(sd, cfg, cuda_device), long_term_mem_gap=meta["gap"])
    tcfg = T.TrainConfig(**meta["train_cfg"])
    n0 = eng.launch_count if hasattr(eng, "launch_count") else 0
    loss, pred, fl, _ = T.train_forward(eng, frames, masks, 1, [n_obj], step=meta["step"], cfg=tcfg)
    torch.cuda.synchronize()
    assert loss.ndim == 0 and loss.is_cuda
    got_fl = [float(x) for x in fl]
    print(f"train_forward on the CUDA engine: loss {float(loss):.5f} (reference {meta['ref_loss']:.5f}), frame losses "
          f"{[round(x, 5) for x in got_fl]} vs {[round(x, 5) for x in meta['ref_frame_losses']]}")
    assert max(abs(a - b) for a, b in zip(got_fl, meta["ref_frame_losses"])) < 3e-3
    assert abs(float(loss) - meta["ref_loss"]) < 5e-3
    got = torch.stack([m[0] for m in pred]).to(torch.uint8).cpu().numpy()
    agree = float((got == z["pred_masks"]).mean())
    print(f"  predicted masks identical to the reference's on {agree:.4%} of the pixels")
    from parity_report import report
    report("training/train_small_forward", vs="the reference's training engine (eval mode, identity shuffle off)",
           loss=float(loss), ref_loss=meta["ref_loss"],
           worst_frame_loss_diff=max(abs(a - b) for a, b in zip(got_fl, meta["ref_frame_losses"])),
           mask_agreement=agree, tolerances={"frame_loss": 3e-3, "loss": 5e-3, "mask": 0.999})
    assert agree >= 0.999
    assert eng.aot_engines[0].long_memories_indexes == meta["ref_idx"]
    # same sample with the previous prediction fed to the memory (use_prev_pred): runs, finite, close to the GT-fed loss
    loss2, _, _, _ = T.train_forward(eng, frames, masks, 1, [n_obj], step=meta["step"], cfg=tcfg, use_prev_pred=True)
    assert torch.isfinite(loss2) and abs(float(loss2) - float(loss)) < 0.5
    assert eng.launch_count > n0
