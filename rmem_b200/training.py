"""Training-side forward of the propagation path (first slice).

Mirrors, for the deterministic part of the reference's training step (paths relative to /root/reference/aot_plus/):

    AOTEngine.forward                       networks/engines/aot_engine.py:40-128   -> train_forward
    AOTEngine.calculate_current_loss        networks/engines/aot_engine.py:484-511  -> LossHead (CUDA, value + gradient)
    AOTEngine._init_losses                  networks/engines/aot_engine.py:130-147  -> TrainConfig, aux_weight
    CrossEntropyLoss top-k schedule         networks/layers/loss.py:190-198         -> top_k_pixels

What is built: the frame sequence of a training sample (reference frame with its auxiliary loss, first propagation, then
memory update with the ground-truth identity -- or with the previous prediction under `use_prev_pred` -- and
propagation for every further frame) through the SAME engine the evaluator uses, and the per-frame loss with its
gradient with respect to the decoder's 1/4-resolution logits as hand-written CUDA (csrc/train_loss.cu).  What is not
built: the backward of the layers under those logits (decoder, GPM / LSTT, encoder), the optimiser and DDP -- the
gradient stops at `grad_logits4`.  The stochastic parts of the reference's training mode (dropout / drop-path, the
random identity shuffle of `restart_engine(batch_size, True)`, aot_engine.py:515-547) are not reproduced: the forward
is the eval-mode arithmetic of the same ops, which is what oracle/make_train_golden.py pins against the reference.
One more eval / train difference matters only for sequences long enough to overflow the bank (the shipped training
samples never are: <= 8 frames at TRAIN_LONG_TERM_MEM_GAP 4 against a capacity of 1 + 8): in training mode the reference
evicts first-in-first-out (restrict_long_memories with use_atten_weight = False, transformer.py:888-890, 964-991), the
engine always applies the inference rule (attention relevance, transformer.py:891-964).

The product path has no CPU fallback: LossHead raises without the CUDA extension.  `train_forward` itself is host
logic over the reference's engine surface and takes the loss as a callable, so the CPU tests drive it with the oracle
engine and the oracle loss.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import torch

from . import _capi

IGNORE = 255
MAX_OBJ = 10


@dataclass
class TrainConfig:
    """The TRAIN_* fields the loss depends on (configs/default.py:45-76; configs/pre_vost.py:13 sets 20000 steps)."""
    total_steps: int = 100000              # TRAIN_TOTAL_STEPS
    top_k_percent_pixels: float = 0.15     # TRAIN_TOP_K_PERCENT_PIXELS
    hard_mining_ratio: float = 0.5         # TRAIN_HARD_MINING_RATIO
    aux_loss_weight: float = 1.0           # TRAIN_AUX_LOSS_WEIGHT
    aux_loss_ratio: float = 1.0            # TRAIN_AUX_LOSS_RATIO


def top_k_pixels(step: int, num_pixels: int, cfg: TrainConfig) -> int:
    """loss.py:167-198: all pixels at step 0, shrinking linearly to top_k_percent of them at hard_mining_ratio *
    total_steps (the 1e-5 added to the mining step is the reference's, loss.py:170)."""
    mining_step = cfg.hard_mining_ratio * cfg.total_steps + 1e-5
    ratio = min(1.0, step / float(mining_step))
    return int((ratio * cfg.top_k_percent_pixels + (1.0 - ratio)) * float(num_pixels))


def aux_weight(step: int, cfg: TrainConfig) -> float:
    """aot_engine.py:53-54 with aux_step of :147."""
    aux_step = cfg.total_steps * cfg.aux_loss_ratio + 1e-5
    return cfg.aux_loss_weight * max(aux_step - step, 0.0) / aux_step


class LossHead:
    """0.5 * bootstrapped cross entropy + 0.5 * soft Jaccard of one frame and its gradient with respect to the 1/4-res
    logits, on the GPU (rmem_train_loss_fwd_bwd).  Returns device tensors; nothing synchronises.  One LossHead owns one
    workspace (histograms, tickets, scratch): use it from one stream at a time -- one head per stream, like one engine per
    stream."""

    def __init__(self, device="cuda:0"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _capi.RmemError("rmem_b200 runs on CUDA devices only (no CPU path)")
        _capi.load()
        self._ws: Optional[torch.Tensor] = None

    def _workspace(self, H: int, W: int) -> torch.Tensor:
        n = C.c_size_t(0)
        _capi.check(_capi.load().rmem_train_loss_workspace_bytes(H, W, C.byref(n)))
        if self._ws is None or self._ws.numel() < n.value:
            self._ws = torch.empty(n.value, dtype=torch.uint8, device=self.device)
        return self._ws

    def __call__(self, logits4: torch.Tensor, gt: torch.Tensor, obj_num: int, top_k: int, want_grad: bool = True,
                 grad_scale: float = 1.0) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """logits4 [1, C, h4, w4] or [C, h4, w4] fp32 (C >= obj_num + 1), gt [.., H, W] integer-valued.
        -> (losses fp32 [3] = total, ce, jaccard; grad like logits4 or None)."""
        lg = logits4.to(self.device, torch.float32).contiguous()
        Cn, h4, w4 = int(lg.shape[-3]), int(lg.shape[-2]), int(lg.shape[-1])
        H, W = int(gt.shape[-2]), int(gt.shape[-1])
        g8 = gt.reshape(H, W).to(self.device).to(torch.uint8).contiguous()
        ws = self._workspace(H, W)
        losses = torch.empty(3, dtype=torch.float32, device=self.device)
        grad = torch.empty_like(lg) if want_grad else None
        with torch.cuda.device(self.device):
            _capi.check(_capi.load().rmem_train_loss_fwd_bwd(
                _capi.ptr(lg), Cn, h4, w4, _capi.ptr(g8), H, W, int(obj_num), C.c_longlong(int(top_k)),
                C.c_float(grad_scale), _capi.ptr(losses), _capi.ptr(grad), _capi.ptr(ws), C.c_size_t(ws.numel()),
                _capi.stream_ptr()))
        return losses, grad


LossFn = Callable[[torch.Tensor, torch.Tensor, int, int], torch.Tensor]     # (logits4, gt [H,W], obj_num, top_k) -> 0-d loss
MaskFn = Callable[[torch.Tensor, int, int, int], torch.Tensor]              # (logits4, H, W, obj_num) -> label [H,W]
# (logits4, gt, obj_num, top_k, scale) -> (0-d loss, scale * d loss / d logits4)
LossGradFn = Callable[[torch.Tensor, torch.Tensor, int, int, float], Tuple[torch.Tensor, torch.Tensor]]


def predict_mask(logits4: torch.Tensor, H: int, W: int, obj_num: int) -> torch.Tensor:
    """predict_current_mask of the training engine (aot_engine.py:467-483; channels above obj_num are at -1e10 there,
    :449-452): uint8 [H, W] = argmax over channels 0..obj_num of the upsampled logits (rmem_train_predict_mask)."""
    lg = logits4.contiguous()
    assert lg.is_cuda and lg.dtype == torch.float32
    Cn, h4, w4 = int(lg.shape[-3]), int(lg.shape[-2]), int(lg.shape[-1])
    lab = torch.empty(H, W, dtype=torch.uint8, device=lg.device)
    with torch.cuda.device(lg.device):
        _capi.check(_capi.load().rmem_train_predict_mask(_capi.ptr(lg), Cn, h4, w4, H, W, int(obj_num), _capi.ptr(lab),
                                                         _capi.stream_ptr()))
    return lab


def _cuda_loss_fn(device) -> LossFn:
    head = LossHead(device)

    def fn(logits4, gt, obj_num, top_k):
        return head(logits4, gt, obj_num, top_k, want_grad=False)[0][0]
    return fn


def _decode_next(engine, img: torch.Tensor):
    """Propagate one frame and leave its 1/4-res logits in the engine (no full-resolution output is needed here)."""
    if hasattr(engine, "propagate_only"):                        # the CUDA engine
        engine.propagate_only(img)
    else:
        engine.match_propogate_one_frame(img, output_size=None)


def _logits4(engine) -> torch.Tensor:
    if hasattr(engine, "logits4_views"):                         # the CUDA engine: a view of the resident logits, no copy
        return engine.logits4_views()[0]
    return engine.aot_engines[0].pred_id_logits


def train_forward(engine, all_frames: torch.Tensor, all_masks: torch.Tensor, batch_size: int, obj_nums: Sequence[int],
                  step: int = 0, tf_board: bool = False, use_prev_pred: bool = False,
                  cfg: Optional[TrainConfig] = None, loss_fn: Optional[LossFn] = None,
                  mask_fn: Optional[MaskFn] = None, logit_grads: Optional[list] = None,
                  loss_grad_fn: Optional[LossGradFn] = None):
    """AOTEngine.forward (aot_engine.py:40-128) over an engine with the reference's inference surface.

    all_frames [(F * B), 3, H, W] and all_masks [(F * B), 1, H, W] are frame-major as the trainer builds them
    (`torch.cat([ref, prev] + curr)`, networks/managers/trainer.py:575-582): frame f of sample b sits at f * B + b.
    The engine runs batch 1, so the samples go through it one after the other.  Returns the reference's tuple
    (loss 0-d, all_pred_mask: F tensors [B, H, W], all_frame_loss: F tensors [B], boards).

    loss_fn / mask_fn default to the CUDA loss head and mask kernel (csrc/train_loss.cu); the CPU tests pass the
    oracle's.

    logit_grads: pass an empty list to also get the head of the backward pass -- it is filled with F tensors
    [B, C, h4, w4], d loss / d pred_id_logits of every frame (the engine overwrites its logits with the next frame, so
    the gradient is taken when the frame's loss is: rmem_train_loss_fwd_bwd with grad_scale = the frame's weight in
    `loss`, w_aux / B for the reference frame and 1 / ((F - 1) B) for the others).  Nothing below the logits is
    differentiated (module docstring)."""
    cfg = cfg or TrainConfig()
    B = int(batch_size)
    assert all_frames.shape[0] % B == 0 and all_frames.shape[0] == all_masks.shape[0], "frame-major [(F*B), ...] inputs"
    F_ = all_frames.shape[0] // B
    assert F_ >= 2, "a training sample is a reference frame plus at least one more frame"
    obj_nums = [int(n) for n in obj_nums]
    assert len(obj_nums) == B and max(obj_nums) <= MAX_OBJ, "one training engine holds up to 10 objects (MODEL_MAX_OBJ_NUM)"
    H, W = int(all_frames.shape[-2]), int(all_frames.shape[-1])
    if loss_fn is None:
        loss_fn = _cuda_loss_fn(all_frames.device if all_frames.is_cuda else "cuda:0")
    if mask_fn is None:
        mask_fn = predict_mask
    w_aux = aux_weight(step, cfg)
    k = top_k_pixels(step, H * W, cfg)
    want_grads = logit_grads is not None
    if want_grads and loss_grad_fn is None:
        head = LossHead(all_frames.device if all_frames.is_cuda else "cuda:0")

        def loss_grad_fn(lg, gt, n, kk, scale):
            losses3, grad = head(lg, gt, n, kk, want_grad=True, grad_scale=scale)
            return losses3[0], grad
    per_sample_grads: List[List[torch.Tensor]] = []

    per_sample_losses: List[List[torch.Tensor]] = []     # [b][f]: aux loss of the reference frame, then the frames' losses
    per_sample_masks: List[List[torch.Tensor]] = []
    for b in range(B):
        n_obj = obj_nums[b]
        losses, masks, grads = [], [], []

        def frame(f, b=b):
            return all_frames[f * B + b: f * B + b + 1]

        def mask(f, b=b):
            return all_masks[f * B + b: f * B + b + 1]

        def loss_and_mask(f, n_obj=n_obj):          # generate_loss_mask (:513-521) on the logits the engine decoded last
            lg = _logits4(engine)
            if want_grads:
                weight = w_aux / B if f == 0 else 1.0 / ((F_ - 1) * B)
                l, g = loss_grad_fn(lg, mask(f)[0, 0], n_obj, k, weight)
                losses.append(l)
                grads.append(g.reshape(g.shape[-3:]))
            else:
                losses.append(loss_fn(lg, mask(f)[0, 0], n_obj, k))
            masks.append(mask_fn(lg, H, W, n_obj))

        engine.restart_engine()
        engine.add_reference_frame(frame(0), mask(0), obj_nums=[n_obj], frame_step=0)       # :68-71
        loss_and_mask(0)                                                                    # :73-79 (auxiliary loss)
        for f in range(1, F_):
            if f > 1:                                                                       # :89-97
                prev = masks[-1].view(1, 1, H, W) if use_prev_pred else mask(f - 1)
                engine.update_memory(prev if prev.dtype == torch.uint8 else prev.float())
            _decode_next(engine, frame(f))                                                  # :82, :98
            loss_and_mask(f)                                                                # :83-86, :99-102
        per_sample_losses.append(losses)
        per_sample_masks.append(masks)
        per_sample_grads.append(grads)

    all_frame_loss = [torch.stack([torch.as_tensor(per_sample_losses[b][f]).float().reshape(()) for b in range(B)])
                      for f in range(F_)]
    all_pred_mask = [torch.stack([per_sample_masks[b][f].reshape(H, W).long() for b in range(B)], dim=0)
                     for f in range(F_)]
    aux_loss = all_frame_loss[0].mean(dim=0)                    # :104 torch.cat of one [B] tensor, mean over it
    pred_loss = torch.cat(all_frame_loss[1:], dim=0).mean(dim=0)    # :105 mean over every frame of every sample
    loss = w_aux * aux_loss + pred_loss                             # :109 (0-d; the trainer's torch.mean is a no-op)
    if want_grads:
        logit_grads[:] = [torch.stack([per_sample_grads[b][f] for b in range(B)], dim=0) for f in range(F_)]
    boards = {"image": {}, "scalar": {}}
    return loss, all_pred_mask, all_frame_loss, boards
