#!/usr/bin/env python
"""Kernel timeline of steady-state c3 frames in the PRODUCTION configuration (prefetched encoder on the side stream, aux
stream inside the layers, PDL on): start / duration / stream of every kernel from CUPTI activity records collected by
torch.profiler (process-wide, so the launches of librmem_b200.so are seen).  ncu serialises launches; this does not.

    python tools/timeline_frame.py --frames 4 > gpurun_out/timeline.txt

Prints one line per kernel (us relative to the first kernel of the window) and per-frame summaries: wall time of the
window, busy time per stream, time during which 1 / 2 / 3 streams had a kernel running.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from rmem_b200.engine import DeAOTModel, RmemConfig, build_engine  # noqa: E402
from rmem_b200.synth import make_state_dict, synthetic_frames, synthetic_label  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--clips", type=int, default=1, help="clips in flight (one engine + stream each)")
    ap.add_argument("--pairs", type=int, default=2, help="frames per encoder pass (prefetch_n, issued that many frames ahead): 2 | 4; 0 = single prefetch")
    ap.add_argument("--json", default="", help="also write the raw records here")
    ap.add_argument("--chrome", default="", help="also export the profiler's chrome trace here")
    a = ap.parse_args()
    H, W, NOBJ, GAP = 481, 849, 10, 5
    dev = torch.device("cuda:0")
    sd = make_state_dict("r50_deaotl", seed=0, sharpen=4.0)
    cfg = RmemConfig(former_mem_len=1, latter_mem_len=7, attn_impl=4, max_engines=1)
    model = DeAOTModel(sd, cfg, dev)
    ring = 8
    engs, srcs, streams = [], [], []
    label0 = synthetic_label(H, W, NOBJ)
    for k in range(a.clips):
        engs.append(build_engine("deaotengine", aot_model=model, long_term_mem_gap=1))
        srcs.append(synthetic_frames(ring + 1, H, W, seed=1000 + k).to(dev))
        streams.append(torch.cuda.Stream(device=dev))

    def step(k, i):
        e, src = engs[k], srcs[k]
        with torch.cuda.stream(streams[k]):
            if a.pairs:
                EG = a.pairs if a.pairs in (2, 4) else 2
                if i % EG == 0:
                    e.prefetch_n([src[1 + (i + EG + j) % ring: 2 + (i + EG + j) % ring] for j in range(EG)])
            else:
                e.prefetch(src[1 + (i + 1) % ring: 2 + (i + 1) % ring])
            lab = e.propagate_label(src[1 + i % ring: 2 + i % ring], output_size=(H, W))
            e.update_memory(lab)

    torch.cuda.synchronize()
    for k in range(a.clips):
        with torch.cuda.stream(streams[k]):
            engs[k].add_reference_frame(srcs[k][0:1], label0.int().to(dev), obj_nums=[NOBJ], frame_step=0)
    i = 0
    for _ in range(10):                       # gap = 1: bank full
        for k in range(a.clips):
            step(k, i)
        i += 1
    for e in engs:
        e.long_term_mem_gap = GAP
    for _ in range(12):
        for k in range(a.clips):
            step(k, i)
        i += 1
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(a.frames):
            for k in range(a.clips):
                step(k, i)
            i += 1
        torch.cuda.synchronize()
    if a.chrome:
        prof.export_chrome_trace(a.chrome)
    evs = []
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA and ev.time_range is not None:
            nm = ev.name
            if nm.startswith("Memcpy") or nm.startswith("Memset"):
                kind = "copy"
            else:
                kind = "kernel"
            evs.append((ev.time_range.start, ev.time_range.end, getattr(ev, "device_resource_id", -1), nm, kind))
    if not evs:
        print("no CUDA activity records (CUPTI unavailable?)")
        return
    evs.sort()
    t0 = evs[0][0]
    import re

    def short(n):
        n = n.replace("(anonymous namespace)::", "").replace("rmem::", "").replace("void ", "")
        return re.sub(r"\(.*", "", n)[:40]
    print(f"# {len(evs)} device activities over {a.frames} frame(s) x {a.clips} clip(s); times in us from the first one")
    print(f"# {'start':>9s} {'dur':>8s} {'stream':>6s}  name")
    for s, e, st, nm, kind in evs:
        print(f"{s - t0:10.1f} {e - s:8.1f} {st:6d}  {short(nm)}")
    wall = evs[-1][1] - t0
    # occupancy of the streams over time
    pts = []
    for s, e, st, nm, kind in evs:
        pts.append((s, 1))
        pts.append((e, -1))
    pts.sort()
    depth, last, hist = 0, pts[0][0], {}
    for t, d in pts:
        hist[depth] = hist.get(depth, 0.0) + (t - last)
        last = t
        depth += d
    per_stream = {}
    for s, e, st, nm, kind in evs:
        per_stream[st] = per_stream.get(st, 0.0) + (e - s)
    nfr = a.frames * a.clips
    print(f"# window {wall:.1f} us = {wall / nfr:.1f} us per frame; sum of activity durations {sum(per_stream.values()) / nfr:.1f} us per frame")
    print("# busy time per stream (us per frame): " + ", ".join(f"s{st}: {v / nfr:.1f}" for st, v in sorted(per_stream.items())))
    print("# concurrently running activities (us per frame): " + ", ".join(f"{k}: {v / nfr:.1f}" for k, v in sorted(hist.items())))
    by = {}
    for s, e, st, nm, kind in evs:
        k = short(nm)
        by.setdefault(k, [0, 0.0])
        by[k][0] += 1
        by[k][1] += e - s
    print("# per kernel (in-stream durations, us per frame):")
    for k, (n, t) in sorted(by.items(), key=lambda kv: -kv[1][1]):
        print(f"#   {k:42s} x{n / nfr:5.1f} {t / nfr:8.1f} us  avg {t / n:6.2f}")
    if a.json:
        import json
        json.dump([dict(start=s - t0, dur=e - s, stream=st, name=nm) for s, e, st, nm, kind in evs], open(a.json, "w"))


if __name__ == "__main__":
    main()
