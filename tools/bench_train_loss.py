"""Time the training-side loss head (csrc/train_loss.cu) at the c3 geometry (481 x 849 label map, 121 x 213 logits, 10
objects) with CUDA events on the launching stream; one JSON line.  The chain is 6 small launches (4 without the gradient) over ~20 MB of scratch:
latency-bound, not bandwidth-bound -- the line says so instead of quoting a roofline fraction for it."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rmem_b200 import synth, training as T  # noqa: E402


def timed(fn, iters=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3          # us


def main():
    dev = torch.device("cuda:0")
    H, W, h4, w4, n_obj = 481, 849, 121, 213, 10
    g = torch.Generator().manual_seed(0)
    lg = (4 * torch.randn(11, h4, w4, generator=g)).to(dev)
    gt = synth.synthetic_label(H, W, n_obj)[0, 0].to(torch.uint8)
    gt[400:440, 600:800] = 255
    gt = gt.to(dev)
    k = T.top_k_pixels(3000, H * W, T.TrainConfig(total_steps=20000))
    head = T.LossHead(dev)
    out = {
        "what": "training loss head, c3 geometry (481x849 labels, 121x213x11 logits, 10 objects), top_k %d of %d" % (k, H * W),
        "fwd_bwd_us": round(timed(lambda: head(lg, gt, n_obj, k)), 2),
        "fwd_only_us": round(timed(lambda: head(lg, gt, n_obj, k, want_grad=False)), 2),
        "predict_mask_us": round(timed(lambda: T.predict_mask(lg, H, W, n_obj)), 2),
        "launches_fwd_bwd": 6, "launches_fwd_only": 4,
        "algorithmic_bytes": {"logits_read": 11 * h4 * w4 * 4, "labels_read": H * W, "grad_written": 11 * h4 * w4 * 4},
        "scratch_bytes": {"ce": H * W * 4, "grad_upsampled": 11 * H * W * 4},
        "bound": "launch latency (6 dependent launches: pixel + 2 histogram passes + top-k sum, each folded by its last block, + 2 gradient kernels); 2.6 MB algorithmic",
        "gpu": torch.cuda.get_device_name(0),
    }
    losses, grad = head(lg, gt, n_obj, k)
    out["losses_total_ce_jaccard"] = [round(float(x), 6) for x in losses]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
