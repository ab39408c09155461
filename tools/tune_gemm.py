#!/usr/bin/env python
"""Tile-shape sweep of the tcgen05 GEMM / implicit-GEMM conv kernel over the shapes ONE steady-state c3 frame launches.

1. rmem_debug_gemm_log records the shape of every gemm_tc launch of a propagated frame (+ memory update, prefetched
   encoder).
2. Every unique shape is rebuilt on synthetic operands (same epilogue features: bias / activation / residual / gate /
   fp32 accumulate / column split) and timed -- CUDA events around `iters` back-to-back launches -- under the launcher's
   own choice and under every forced (BN, split-K, stages) of rmem_debug_gemm_force.
Prints one JSON line per (shape, config) and a summary of what a per-shape best choice would save per frame.

    python tools/tune_gemm.py > gpurun_out/tune_gemm.jsonl
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from rmem_b200 import _capi  # noqa: E402
from rmem_b200.engine import DeAOTModel, RmemConfig, build_engine  # noqa: E402
from rmem_b200.synth import make_state_dict, synthetic_frames, synthetic_label  # noqa: E402

FIELDS = ["M", "N", "K", "conv", "Hin", "Win", "Cin", "Wout", "kw", "stride", "pad", "act", "res", "gate", "flags", "n_split"]


def record_frame_shapes(dev, H=481, W=849, objects=10, latter=7):
    lib = _capi.load()
    sd = make_state_dict("r50_deaotl", seed=0, sharpen=4.0)
    cfg = RmemConfig(former_mem_len=1, latter_mem_len=latter, attn_impl=4, max_engines=1)
    eng = build_engine("deaotengine", aot_model=DeAOTModel(sd, cfg, dev), long_term_mem_gap=1)
    frames = synthetic_frames(4, H, W, seed=1000).to(dev)
    label0 = synthetic_label(H, W, objects)
    eng.add_reference_frame(frames[0:1], label0.int().to(dev), obj_nums=[objects], frame_step=0)
    for i in range(latter + 3):
        lab = eng.propagate_label(frames[1 + i % 3:2 + i % 3], output_size=(H, W))
        eng.update_memory(lab)
    eng.long_term_mem_gap = 5
    torch.cuda.synchronize()
    cap = 512
    buf = (C.c_int * (cap * 16))()
    _capi.check(lib.rmem_debug_gemm_log(buf, cap))
    lab = eng.propagate_label(frames[1:2], output_size=(H, W))      # a frame that does not append to the bank
    eng.update_memory(lab)
    torch.cuda.synchronize()
    n = lib.rmem_debug_gemm_log_count()
    _capi.check(lib.rmem_debug_gemm_log(None, 0))
    recs = [tuple(buf[i * 16:(i + 1) * 16]) for i in range(n)]
    del eng
    return recs


def build_desc(rec, dev, g):
    r = dict(zip(FIELDS, rec))
    OP = _capi.op_dtype()
    M, N, K = r["M"], r["N"], r["K"]
    keep = []
    d = _capi.GemmDesc()
    if r["conv"] == 2:                      # stem: padded NHWC8 image, K = kw * 64
        A = torch.randn(r["Hin"], r["Win"], 8, generator=g).to(dev).to(OP)
        d.A = A.data_ptr()
    elif r["conv"] == 1:
        A = torch.randn(r["Hin"], r["Win"], r["Cin"], generator=g).to(dev).to(OP)
        d.A = A.data_ptr()
    else:
        A = torch.randn(M, K, generator=g).to(dev).to(OP)
        d.A, d.lda = A.data_ptr(), K
    Bw = (torch.randn(N, K, generator=g) * 0.05).to(dev).to(OP)
    d.B, d.ldb = Bw.data_ptr(), K
    d.M, d.N, d.K = M, N, K
    d.conv, d.Hin, d.Win, d.Cin, d.Wout, d.kw, d.stride, d.pad = (r["conv"], r["Hin"], r["Win"], r["Cin"], r["Wout"],
                                                                  r["kw"], r["stride"], r["pad"])
    d.alpha = 1.0
    bias_m = (r["flags"] >> 2) & 1
    bias = torch.randn(M if bias_m else ((N + 31) // 32 * 32), generator=g).to(dev)
    d.bias, d.bias_along_m = bias.data_ptr(), bias_m
    d.act, d.act_from_col = r["act"], 0
    Np = (N + 31) // 32 * 32
    keep += [A, Bw, bias]
    if r["res"]:
        res = torch.randn(M, Np, generator=g).to(dev).to(OP)
        d.residual, d.ldr = res.data_ptr(), Np
        keep.append(res)
    if r["gate"]:
        gt = torch.randn(M, Np, generator=g).to(dev).to(OP)
        d.gate, d.ldg = gt.data_ptr(), Np
        keep.append(gt)
    f32 = r["flags"] & 1
    d.accumulate = (r["flags"] >> 1) & 1
    ns = r["n_split"]
    if ns:
        c1 = torch.zeros(M, ns, dtype=torch.float32 if f32 else OP, device=dev)
        c2 = torch.zeros(M, Np - ns, dtype=OP, device=dev)
        d.C, d.ldc, d.c_is_f32 = c1.data_ptr(), ns, f32
        d.C2, d.ldc2, d.c2_is_f32, d.n_split = c2.data_ptr(), Np - ns, 0, ns
        keep += [c1, c2]
    else:
        c1 = torch.zeros(M, Np, dtype=torch.float32 if f32 else OP, device=dev)
        d.C, d.ldc, d.c_is_f32 = c1.data_ptr(), Np, f32
        d.n_split = 1 << 30
        keep.append(c1)
    d.pad_n_ok = int(N % 32 != 0)
    return d, keep


def time_desc(lib, d, iters):
    st = _capi.stream_ptr()
    for _ in range(3):
        _capi.check(lib.rmem_gemm_fwd(C.byref(d), st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        _capi.check(lib.rmem_gemm_fwd(C.byref(d), st))
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=40)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    lib = _capi.load()
    recs = record_frame_shapes(dev)
    counts = {}
    for r in recs:
        counts[r] = counts.get(r, 0) + 1
    print(f"# {len(recs)} gemm_tc launches per frame, {len(counts)} unique shapes", file=sys.stderr)
    g = torch.Generator().manual_seed(0)
    total_auto = total_best = 0.0
    summary = []
    for rec, cnt in counts.items():
        r = dict(zip(FIELDS, rec))
        d, keep = build_desc(rec, dev, g)
        nk = r["K"] // 64
        _capi.check(lib.rmem_debug_gemm_force(0, 0, 0))
        t_auto = time_desc(lib, d, a.iters)
        cands = []
        for bn in (64, 128, 256):
            if bn > 64 and r["N"] < bn:
                continue
            for s in ((1, 2, 3, 4) if bn == 64 else (1,)):
                if s > 1 and s > nk // 2:
                    continue
                for stg in (0, 2, 6 if bn != 256 else 4):
                    if stg and stg > max(2, (nk + s - 1) // s):
                        continue
                    cands.append((bn, s, stg))
        best = (t_auto, (0, 0, 0))
        for bn, s, stg in cands:
            _capi.check(lib.rmem_debug_gemm_force(bn, s, stg))
            try:
                t = time_desc(lib, d, a.iters)
            except Exception as ex:           # a forced shape the kernel refuses: skip
                print(json.dumps({"shape": r, "force": [bn, s, stg], "error": str(ex)[:200]}))
                torch.cuda.synchronize()
                continue
            print(json.dumps({"shape": r, "count": cnt, "force": [bn, s, stg], "us": round(t, 2), "auto_us": round(t_auto, 2)}))
            if t < best[0]:
                best = (t, (bn, s, stg))
        _capi.check(lib.rmem_debug_gemm_force(0, 0, 0))
        total_auto += cnt * t_auto
        total_best += cnt * best[0]
        summary.append((cnt * (t_auto - best[0]), cnt, r, t_auto, best))
        del keep
    print("# per-shape summary (sorted by saving per frame):")
    for save, cnt, r, t_auto, best in sorted(summary, key=lambda x: -x[0]):
        tag = f"conv{r['kw']}x{r['kw']}s{r['stride']}" if r["conv"] else "linear"
        print(f"# x{cnt:2d} M={r['M']:6d} N={r['N']:5d} K={r['K']:5d} {tag:10s} res={r['res']} gate={r['gate']} fl={r['flags']} "
              f"auto {t_auto:7.2f} us  best {best[0]:7.2f} us @ BN,S,stages={best[1]}  saves {save:6.1f} us/frame")
    print(f"# total per frame: auto {total_auto:.1f} us -> per-shape best {total_best:.1f} us")


if __name__ == "__main__":
    main()
