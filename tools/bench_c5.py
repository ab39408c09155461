#!/usr/bin/env python
"""BASELINE.json configs[4] ("c5"): a batch of 2000-frame 480p clips (R50_DeAOTL + RMem, 10 objects, T=8) sharded per clip
over the GPUs of one node through the evaluator shell: every rank pulls the next clip from ONE atomic counter in the
torch.distributed store (the reference's dynamic clip queue, evaluator.py:276-295, tools/eval.py:137-145), runs the
reference's per-clip loop (rmem_b200.evaluator.evaluate_clip: long-term gap = max(round(2000/30), 5) = 67,
evaluator.py:330-332) and (frames, seconds) are gathered once at the end -- no collective on the per-frame path.

    python tools/bench_c5.py --clips 4                                            # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \\
        tools/bench_c5.py --clips 32

End to end by construction: every frame starts in pinned HOST memory (a ring of distinct synthetic frames per clip, clip i
seeded 1000 + i) and every uint8 label map ends in pinned host memory; both copies are inside the timed region.
Prints ONE JSON line on rank 0: all-frame FPS the reference's way (sum of frames / sum of per-frame CUDA-event seconds over
all ranks, evaluator.py:589-613), wall-clock job throughput (frames of all ranks / slowest rank's wall time between two
barriers) and the per-rank imbalance."""
import argparse
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

H, W, N_OBJ = 481, 849, 10


class SyntheticClip:
    """A clip of `n_frames` 480p frames cycling through a pinned ring of `ring` distinct frames; frame 0 carries the
    label of its 10 objects.  Same sample dict as rmem_b200.evaluator.ClipDataset."""

    def __init__(self, index: int, n_frames: int, ring: int = 8):
        from rmem_b200.synth import synthetic_frames, synthetic_label
        self.seq_name = f"synthetic_{index:04d}"
        self.n = n_frames
        self.frames = synthetic_frames(ring + 1, H, W, seed=1000 + index).pin_memory()
        self.label0 = synthetic_label(H, W, N_OBJ).int()
        self.ring = ring

    def __len__(self):
        return self.n

    def __getitem__(self, idx):
        j = 0 if idx == 0 else 1 + (idx - 1) % self.ring
        s = {"current_img": self.frames[j:j + 1],
             "meta": {"seq_name": self.seq_name, "frame_num": self.n, "obj_num": N_OBJ, "current_name": f"{idx:05d}.jpg",
                      "height": H, "width": W, "flip": False, "obj_idx": list(range(N_OBJ + 1))}}
        if idx == 0:
            s["current_label"] = self.label0
        return s


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=4)
    ap.add_argument("--frames", type=int, default=2000)
    ap.add_argument("--static", action="store_true", help="static round-robin instead of the shared counter")
    ap.add_argument("--in-flight", type=int, default=1,
                    help="clips processed concurrently per GPU (one engine + stream + host thread each): clips are "
                         "independent, and a single clip's chain of small kernels leaves most SMs idle")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    store = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)
        if not a.static:
            store = dist.distributed_c10d._get_default_store()
    from rmem_b200.engine import DeAOTModel, RmemConfig, build_engine
    from rmem_b200.evaluator import evaluate_clips, long_term_gap
    from rmem_b200.sharding import broadcast_weights
    from rmem_b200.synth import make_state_dict
    sd = make_state_dict("r50_deaotl", seed=0, sharpen=4.0) if rank == 0 else None
    sd = broadcast_weights(sd, dev, world)
    model = DeAOTModel(sd, RmemConfig(former_mem_len=1, latter_mem_len=7, max_engines=1), dev)
    engines = [build_engine("deaotengine", aot_model=model) for _ in range(max(1, a.in_flight))]
    eng = engines[0]
    clips = [SyntheticClip(i, a.frames) for i in range(a.clips)] if world == 1 else None
    if world > 1:
        # every rank may draw any clip from the queue: build them lazily to keep host memory bounded
        class Lazy:
            def __len__(self_inner):
                return a.clips

            def __getitem__(self_inner, i):
                return SyntheticClip(i, a.frames)
        clips = Lazy()
    # warm-up clip (not counted): first-use costs (tensor maps, lazy allocations)
    from rmem_b200.evaluator import evaluate_clip
    evaluate_clip(eng, SyntheticClip(10_000 + rank, 120), device=dev)
    # several clips in flight on this GPU: one engine + host thread + CUDA stream per clip, all drawing from the same queue
    # (rmem_b200.evaluator.evaluate_clips, extra_engines)
    for e in engines[1:]:                                   # per-engine warm-up (lazy allocations, tensor maps)
        evaluate_clip(e, SyntheticClip(20_000 + rank * 8 + len(engines), 60), device=dev)
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    t0 = time.perf_counter()
    out = evaluate_clips(eng, clips, device=dev, rank=rank, world=world, store=store, log=None, extra_engines=engines[1:])
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    walls = [wall]
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([wall], dtype=torch.float64, device=dev)
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        walls = [float(o.item()) for o in outs]
        dist.barrier()
    if rank == 0:
        per_rank = out["per_rank"]
        frames_total = sum(f for f, _ in per_rank)
        rec = {"metric": "VOS frames/sec, 2000-frame 480p clips sharded per clip (c5)", "unit": "frames/s", "n_gpus": world,
               "value": round(frames_total / max(walls), 2),
               "value_is": "all propagated frames of the job / slowest rank's wall time (barrier to barrier), host frames in, "
                           "host labels out",
               "all_frame_fps_reference_style": round(out["all_frame_fps"], 2),
               "all_frame_fps_note": "sum of frames / sum of per-frame CUDA-event seconds over ranks (evaluator.py:589-613): a "
                                     "per-GPU rate, multiply by n_gpus for the job",
               "clips": a.clips, "frames_per_clip": a.frames, "gap": long_term_gap(a.frames), "clips_in_flight_per_gpu": a.in_flight,
               "queue": "static round-robin" if (a.static or world == 1) else "shared atomic counter in the c10d store",
               "per_rank_frames": [f for f, _ in per_rank], "per_rank_event_seconds": [round(s, 3) for _, s in per_rank],
               "per_rank_wall_seconds": [round(w, 3) for w in walls],
               "imbalance": round(max(walls) / (sum(walls) / len(walls)), 4),
               "e2e": {"h2d_bytes_per_frame": 3 * H * W * 4, "d2h_bytes_per_frame": H * W}}
        print(json.dumps(rec), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
