"""Long / deep-bank golden fixtures from the CPU oracle  --  TEST INFRASTRUCTURE (runs in the build container, CPU only).

    python oracle/make_long_golden.py [case ...]

The GPU parity tests of BASELINE.json's full-size configs cannot afford the fp32 CPU oracle at test time for hundreds of
frames (1.4 s per 480p frame, ~25 s per 720p / 30-object frame), so the oracle is run HERE once and its observables are
committed as small fixtures (tests/golden/long_*.npz); tests/test_long_clips_gpu.py replays the same clip through the CUDA
engine and compares.  The oracle itself is pinned against the unmodified reference by oracle/make_golden.py.

Both sides are TEACHER-FORCED with the same procedural label history (the reference clip's rectangles, rolled by a few
pixels per frame), so nothing but seeds has to be stored for the inputs: weights, frames and labels are regenerated from
(seed, frame index).  Stored per clip: long_memories_indexes after every update, the normalised relevance vector and
the dropped position of every eviction, and -- on a strided subset of frames -- a strided sample of the 1/4-resolution
logits plus the oracle's predicted label map (zlib-compressed).

Cases: c3 (R50_DeAOTL+RMem, 481x849, 10 objects, T=8) for 320 frames at gap 5 and at gap 67 = max(round(2000/30), 5)
(evaluator.py:330-332); c3 at the shipped bank capacity T=9; c4 (721x1281, 30 objects = 3 engines) at T=8.
"""
from __future__ import annotations

import os
import sys
import time
import zlib

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import rmem_oracle as O  # noqa: E402

CASES = {
    # name: (H, W, n_obj, former, latter, gap, n_frames, logit_every, seed)
    "long_c3_gap5": (481, 849, 10, 1, 7, 5, 321, 16, 21),
    "long_c3_gap67": (481, 849, 10, 1, 7, 67, 321, 16, 22),
    "deep_c3_T9": (481, 849, 10, 1, 8, 1, 14, 1, 23),
    "deep_c4_T8": (721, 1281, 30, 1, 7, 1, 12, 2, 24),
}
RING = 8          # distinct frames per clip (cycled, like bench.py)
SAMPLE = 4        # logits are stored at every SAMPLE-th row / column of the 1/4-resolution map


def forced_label(label0: torch.Tensor, f: int) -> torch.Tensor:
    """The label history both implementations are fed: the reference rectangles drifting by (2, 3) pixels per frame."""
    return torch.roll(label0, shifts=(2 * f % 37, 3 * f % 53), dims=(2, 3))


def clip_inputs(name):
    H, W, n_obj, former, latter, gap, n_frames, every, seed = CASES[name]
    frames = O.synthetic_frames(RING + 1, H, W, seed=seed)
    label0 = O.synthetic_label(H, W, n_obj)
    return frames, label0


def frame_of(frames, f):
    return frames[0:1] if f == 0 else frames[1 + (f - 1) % RING: 2 + (f - 1) % RING]


def run(name: str):
    H, W, n_obj, former, latter, gap, n_frames, every, seed = CASES[name]
    torch.manual_seed(0)
    sd = O.make_state_dict("r50_deaotl", seed=seed, sharpen=4.0)
    frames, label0 = clip_inputs(name)
    eng = O.OracleEngine(sd, O.OracleConfig(former_mem_len=former, latter_mem_len=latter), long_term_mem_gap=gap)
    idx_hist, evict_frames, evict_rel, evict_drop, logit_frames, logit_samples, label_blobs = [], [], [], [], [], [], []
    t0 = time.time()
    with torch.no_grad():
        eng.restart_engine()
        eng.add_reference_frame(frames[0:1], label0, obj_nums=[n_obj], frame_step=0)
        for f in range(1, n_frames):
            lg = eng.match_propogate_one_frame(frame_of(frames, f), output_size=(H, W))
            if f % every == 0 or f == n_frames - 1:
                logit_frames.append(f)
                logit_samples.append(np.stack([e.pred_id_logits[0, :, ::SAMPLE, ::SAMPLE].numpy().astype(np.float32)
                                               for e in eng.aot_engines]))
                lab = O.logits_to_label(lg)[0, 0].numpy().astype(np.uint8)
                label_blobs.append(np.frombuffer(zlib.compress(lab.tobytes(), 9), dtype=np.uint8))
            for e in eng.aot_engines:
                e.last_drop = None
            eng.update_memory(forced_label(label0, f))
            idx_hist.append([list(e.long_memories_indexes) for e in eng.aot_engines])
            for gi, e in enumerate(eng.aot_engines):
                if getattr(e, "last_drop", None) is not None:
                    evict_frames.append((f, gi))
                    evict_rel.append(e.last_rel.numpy().astype(np.float32))
                    evict_drop.append(int(e.last_drop))
            if f % 20 == 0:
                print(f"[{name}] frame {f}/{n_frames - 1}  {time.time() - t0:.0f}s  bank {idx_hist[-1][0]}", flush=True)
    n_eng = len(eng.aot_engines)
    cap = former + latter
    idx_arr = np.full((len(idx_hist), n_eng, cap + 1), -1, dtype=np.int32)
    for i, per in enumerate(idx_hist):
        for gi, l in enumerate(per):
            idx_arr[i, gi, :len(l)] = l
    rel_arr = np.zeros((len(evict_rel), cap + 1), dtype=np.float32)
    for i, r in enumerate(evict_rel):
        rel_arr[i, :len(r)] = r
    out = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(out, case=np.array(CASES[name]), idx=idx_arr, evict_frames=np.array(evict_frames, dtype=np.int32).reshape(-1, 2),
                        evict_rel=rel_arr, evict_drop=np.array(evict_drop, dtype=np.int32),
                        logit_frames=np.array(logit_frames, dtype=np.int32), logits=np.stack(logit_samples).astype(np.float16),
                        labels_blob=np.concatenate(label_blobs),
                        labels_off=np.cumsum([0] + [len(b) for b in label_blobs]).astype(np.int64))
    print(f"[{name}] wrote {out}: {len(evict_drop)} evictions, {len(logit_frames)} logit samples, "
          f"{os.path.getsize(out) / 1e6:.2f} MB, {time.time() - t0:.0f}s", flush=True)


if __name__ == "__main__":
    torch.set_num_threads(int(os.environ.get("ORACLE_THREADS", "4")))
    for nm in (sys.argv[1:] or list(CASES)):
        run(nm)
