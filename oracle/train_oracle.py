"""CPU restatement of the training-side loss of the reference (first slice).

TEST INFRASTRUCTURE ONLY -- imported by tests/ and oracle/make_train_golden.py, never by rmem_b200/.  Plain torch fp32
on CPU, differentiable through autograd, so that the same function is the oracle of the forward value AND of the
gradient the CUDA loss head returns.  Pinned against the UNMODIFIED reference (its `CrossEntropyLoss` /
`SoftJaccordLoss` modules and `AOTEngine.calculate_current_loss`, run in the build container) by
oracle/make_train_golden.py -> tests/golden/train_small.npz.

Restates (paths relative to /root/reference/aot_plus/):
    AOTEngine.calculate_current_loss        networks/engines/aot_engine.py:484-511
    CrossEntropyLoss.forward (top-k branch) networks/layers/loss.py:163-211
    SoftJaccordLoss.forward + tversky_loss  networks/layers/loss.py:30-56, 136-160  (alpha = beta = 1, eps 1e-6)
    flatten_probas (ignore = 255)           networks/layers/loss.py:59-74
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

IGNORE = 255


def loss_head(logits4: Tensor, gt: Tensor, obj_num: int, top_k: int) -> Tuple[Tensor, Tensor, Tensor]:
    """logits4 [1, >=obj_num+1, h4, w4] fp32 (the engine's pred_id_logits), gt [H, W] integer-valued (255 = ignore).
    Returns (0.5 * ce + 0.5 * jaccard, ce, jaccard) as 0-d tensors (aot_engine.py:141-142: both weights 0.5)."""
    H, W = int(gt.shape[-2]), int(gt.shape[-1])
    n_ch = obj_num + 1
    up = F.interpolate(logits4, size=(H, W), mode="bilinear", align_corners=True)[0, :n_ch]   # :487-492, :499
    x = up.reshape(n_ch, H * W).t()                                                           # [P, C]
    lab = gt.reshape(-1).long()
    valid = lab != IGNORE
    # bootstrapped cross entropy: mean of the top_k largest per-pixel losses; ignored pixels count as zeros among
    # the H*W candidates (reduction='none' with ignore_index, loss.py:176-177, 199-203)
    pix = F.cross_entropy(x, lab, ignore_index=IGNORE, reduction="none")
    ce = torch.topk(pix, k=top_k).values.mean()
    # soft Jaccard over the classes that own at least one valid pixel
    prob = torch.softmax(x, dim=1)[valid]
    lv = lab[valid]
    terms = []
    for c in range(n_ch):
        fg = (lv == c).to(prob.dtype)
        if float(fg.sum()) == 0:
            continue
        p = prob[:, c]
        inter = (p * fg).sum()
        denom = inter + (p * (1 - fg)).sum() + ((1 - p) * fg).sum()
        terms.append(1 - inter / (denom + 1e-6))
    jac = torch.stack(terms).mean() if terms else prob.sum() * 0
    return 0.5 * ce + 0.5 * jac, ce, jac


def loss_head_with_grad(logits4: Tensor, gt: Tensor, obj_num: int, top_k: int):
    """(total, ce, jaccard, d total / d logits4) -- the four outputs of rmem_train_loss_fwd_bwd."""
    x = logits4.detach().clone().requires_grad_(True)
    total, ce, jac = loss_head(x, gt, obj_num, top_k)
    total.backward()
    return float(total.detach()), float(ce.detach()), float(jac.detach()), x.grad.detach()


def predict_mask(logits4: Tensor, H: int, W: int, obj_num: int) -> Tensor:
    """predict_current_mask of the training engine (aot_engine.py:467-483 after :449-452): [H, W] long."""
    up = F.interpolate(logits4.reshape(1, *logits4.shape[-3:]), size=(H, W), mode="bilinear", align_corners=True)
    return torch.argmax(up[0, :obj_num + 1], dim=0)
