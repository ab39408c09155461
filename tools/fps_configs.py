#!/usr/bin/env python
"""Steady-state frames/s of the other BASELINE.json configs (parity-test cases, not bench lines): c2 R50_AOTL+RMem 480p
1 object T=4, c3 (the bench workload), c4 R50_DeAOTL+RMem 720p 30 objects (3 object groups) T=8.  Device-resident
frames, next-frame prefetch on, CUDA events over `--frames` frames after the bank is full."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from rmem_b200.engine import RmemModel, RmemConfig, build_engine
from rmem_b200.synth import make_state_dict, synthetic_frames, synthetic_label

CONFIGS = {"c2": ("r50_aotl", 481, 849, 1, 3), "c3": ("r50_deaotl", 481, 849, 10, 7), "c4": ("r50_deaotl", 721, 1281, 30, 7)}

def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--frames", type=int, default=60); a = ap.parse_args()
    dev = torch.device("cuda:0")
    for name, (model, H, W, n_obj, latter) in CONFIGS.items():
        sd = make_state_dict(model, seed=0, sharpen=4.0)
        cfg = RmemConfig(model=model, former_mem_len=1, latter_mem_len=latter, max_engines=(n_obj + 9) // 10)
        eng = build_engine("deaotengine" if model == "r50_deaotl" else "aotengine", aot_model=RmemModel(sd, cfg, dev),
                           long_term_mem_gap=1)
        frames = synthetic_frames(9, H, W, seed=1000).to(dev)
        R = 8
        EG = int(os.environ.get("RMEM_BENCH_ENC_GROUP", "2"))
        pairs = EG > 1
        eng.add_reference_frame(frames[0:1], synthetic_label(H, W, n_obj).int().to(dev), obj_nums=[n_obj], frame_step=0)
        def step(i):
            if pairs:
                if i % EG == 0:
                    eng.prefetch_n([frames[1 + (i + EG + j) % R:2 + (i + EG + j) % R] for j in range(EG)])
            else:
                eng.prefetch(frames[1 + (i + 1) % R:2 + (i + 1) % R])
            lab = eng.propagate_label(frames[1 + i % R:2 + i % R], output_size=(H, W))
            eng.update_memory(lab)
        n0 = latter + 3 + (-(latter + 3)) % 4
        for i in range(n0):
            step(i)
        eng.long_term_mem_gap = 5
        for i in range(n0, n0 + 10):
            step(i)
        n0 += 10
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n0, n0 + a.frames):
            step(i)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.frames
        T = len(eng.aot_engines[0].long_memories_indexes)
        print(f"{name}: {model} {H}x{W} {n_obj} objects ({len(eng.aot_engines)} groups) T={T}: {ms:.3f} ms/frame = {1e3 / ms:.1f} frames/s", flush=True)
        del eng
        torch.cuda.empty_cache()

if __name__ == "__main__":
    main()
