#!/usr/bin/env python
"""Headline benchmark: VOS frames/sec @480p, R50_DeAOTL + RMem, T=8 memory bank, 10 objects (BASELINE.json
configs[2] = "c3"), one process per GPU, independent clips sharded per rank (no data-path collective).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
    python bench.py --impl reference ...                     # CPU reference arm (the oracle port of aot_plus)

A "step" is one propagated frame = match_propogate_one_frame + mask-ID assignment + update_memory -- the
reference's own FPS bracket (aot_plus/networks/managers/evaluator.py:399-404, 525-527).  `value` times the steps
with frames already resident in HBM; `e2e` times the same call sequence from pinned host frames with the
host->device copy of every frame and the device->host read of every uint8 label map inside the timed region.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

WORKLOAD = "c3: R50_DeAOTL+RMem, 480p (481x849 -> 1674 tokens), 10 objects, T=8 (former 1 + latter 7), gap 5"
H, W, N_OBJ, FORMER, LATTER, GAP = 481, 849, 10, 1, 7, 5
METRIC = "VOS frames/sec @480p R50_DeAOTL+RMem T=8"
# long-term attention algorithmic FLOPs per layer at c3 = 2*HW*(T*HW)*(Dk+Dv)
HW_TOK = 31 * 54
LT_FLOPS_PER_LAUNCH = 2.0 * HW_TOK * (8 * HW_TOK) * (128 + 1024)
ATTN_IMPLS = {"dense": 0, "tc2": 2, "tc3": 3, "tc4": 4}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, gpu_index: int):
        self.samples = []
        self.reasons = set()
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.t = threading.Thread(target=self._read, daemon=True)
        self.t.start()

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                self.samples.append((float(parts[0]), float(parts[1])))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    self.reasons.add(n)

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        if not self.samples:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=sorted(self.reasons))
        sm = sorted(s[0] for s in self.samples)
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(s[1] for s in self.samples), reasons=sorted(self.reasons),
                    samples=len(sm))


PREFETCH = os.environ.get("RMEM_BENCH_PREFETCH", "1") != "0"


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG", "WARN")      # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    from rmem_b200 import _capi
    from rmem_b200.engine import DeAOTModel, RmemConfig, build_engine
    from rmem_b200.synth import make_state_dict, synthetic_frames, synthetic_label
    from rmem_b200.sharding import broadcast_weights

    sd = make_state_dict("r50_deaotl", seed=0, sharpen=4.0) if rank == 0 else None
    sd = broadcast_weights(sd, dev, world)           # one NCCL broadcast of the weights at init (north_star)
    cfg = RmemConfig(former_mem_len=FORMER, latter_mem_len=LATTER,
                     attn_impl=ATTN_IMPLS[args.attn], max_engines=1)
    eng = build_engine("deaotengine", aot_model=DeAOTModel(sd, cfg, dev), long_term_mem_gap=GAP)

    # independent clip per rank (clip i seeded 1000+i); frames cycle through a resident ring
    ring = 8
    frames = synthetic_frames(ring + 1, H, W, seed=1000 + rank)
    label0 = synthetic_label(H, W, N_OBJ)
    frames_dev = frames.to(dev)
    frames_pin = frames.pin_memory()
    fill = (FORMER + LATTER + 1) * GAP + 2            # frames until the bank is full (T = 8 steady)

    def clip_start(src):
        eng.restart_engine()
        eng.long_term_mem_gap = GAP
        eng.add_reference_frame(src[0:1], label0.int().to(dev), obj_nums=[N_OBJ], frame_step=0)

    EG = args.enc_group if PREFETCH else 1               # frames per encoder pass (1 = single-frame prefetch)
    PAIRS = EG > 1

    def fr(src, k):
        return src[1 + k % ring: 2 + k % ring]

    def step(i, src):
        # software pipelining across frames: the image encoder runs on the engine's side stream while frame i propagates
        # -- frame i+1 alone, or (pair mode) frames i+2 and i+3 in one pass every second frame
        if PAIRS:
            if i % EG == 0:
                eng.prefetch_n([fr(src, i + EG + j) for j in range(EG)])
        elif PREFETCH:
            eng.prefetch(fr(src, i + 1))
        lab = eng.propagate_label(fr(src, i), output_size=(H, W))
        eng.update_memory(lab)        # 480p -> output size == input size, nearest resize is the identity
        return lab

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # End-to-end leg: every frame starts in PINNED HOST memory and every label map ends in pinned host memory.  The
    # host->device copy of frame i+1 is issued on a copy stream while frame i computes (what a prefetching loader does,
    # evaluator.py:308,372 uses pin_memory + non_blocking); both copies are inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    d2h_stream = torch.cuda.Stream(device=dev)             # label maps leave on their own stream: a device->host copy in the
    lab_ready = [torch.cuda.Event() for _ in range(4)]     # caller's stream would sit between two frames of the chain (~20 us)
    NS = 3 * EG if PAIRS else 2                            # staging buffers: frames i-1 .. i+2*EG-1 are live in group mode
    stage = [torch.empty(1, 3, H, W, dtype=torch.float32, device=dev) for _ in range(NS)]
    staged = [torch.cuda.Event() for _ in range(NS)]
    consumed = [torch.cuda.Event() for _ in range(NS)]
    counter = {"dev": 0, "e2e": 0}                         # frame index runs on across the timed blocks

    E2E_DEBUG = os.environ.get("RMEM_BENCH_E2E_DEBUG", "")   # "noh2d" / "nod2h": leave one copy out (where does the e2e gap come from?)

    def prefetch(i, src):
        b = i % NS
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[b])           # frame i-NS has finished reading this staging buffer
            if E2E_DEBUG != "noh2d" or i < 2 * NS:
                stage[b].copy_(fr(src, i), non_blocking=True)
            staged[b].record(copy_stream)

    def timed(src, steps, e2e):
        host_lab = torch.empty(2, 1, 1, H, W, dtype=torch.uint8).pin_memory()
        main = torch.cuda.current_stream()
        first = counter["e2e" if e2e else "dev"]
        if e2e and first == 0:
            for b in range(NS):
                consumed[b].record(main)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if e2e and first == 0:                              # very first e2e frame(s): nothing staged yet
            for k in range(2 * EG if PAIRS else 1):
                prefetch(k, src)
        for i in range(first, first + steps):
            if e2e:
                # The copy of a frame and its encoding are issued ahead of its propagate: frame i+1 (single mode), or
                # frames i+2, i+3 every second frame (pair mode).  The last one or two frames staged by a block are
                # consumed by the next block (the frame index runs on), so every block copies and encodes exactly
                # `steps` frames inside its timed region.
                if PAIRS:
                    # the copies of a group are issued one step before its encoder pass, so that the pass (which has to
                    # be done two steps later) does not start by waiting ~0.2 ms for PCIe
                    if i % EG == 0:
                        eng.prefetch_n([stage[(i + EG + j) % NS] for j in range(EG)], stream=copy_stream)
                    if (i + 1) % EG == 0:
                        for j in range(EG):
                            prefetch(i + 1 + EG + j, src)
                else:
                    prefetch(i + 1, src)
                    if PREFETCH:
                        eng.prefetch(stage[(i + 1) % NS], stream=copy_stream)   # encode i+1 once its copy has landed
                main.wait_event(staged[i % NS])
                lab = eng.propagate_label(stage[i % NS], output_size=(H, W))
                consumed[i % NS].record(main)
                lab_ready[i % 4].record(main)
                eng.update_memory(lab)
                if E2E_DEBUG != "nod2h":
                    with torch.cuda.stream(d2h_stream):
                        d2h_stream.wait_event(lab_ready[i % 4])
                        host_lab[i % 2].copy_(lab, non_blocking=True)
                        lab.record_stream(d2h_stream)
            else:
                lab = step(i, src)
        counter["e2e" if e2e else "dev"] = first + steps
        if e2e:
            main.wait_stream(d2h_stream)                    # every label map has landed before the region ends
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        print(f"[bench] rank {rank}/{world} {'e2e' if e2e else 'device'}: {ms / steps:.4f} ms/step", file=sys.stderr, flush=True)
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # RMEM_BENCH_HIPRIO=1: run the propagation path on a high-priority stream (measured: worse, the overlap with the
    # prefetched encoder is lost -- 1.52 vs 1.335 ms per frame; kept as a switch).
    hp = torch.cuda.Stream(device=dev, priority=-1) if os.environ.get("RMEM_BENCH_HIPRIO", "0") != "0" else None
    if hp is not None:
        hp.wait_stream(torch.cuda.current_stream())
        torch.cuda.set_stream(hp)
    # ---- device-resident run ----
    clip_start(frames_dev)
    nwarm = max(fill, args.warmup)
    nwarm += (-nwarm) % 4                                 # group mode: blocks start on a multiple of the group size
    for i in range(nwarm):
        step(i, frames_dev)
    counter["dev"] = nwarm
    assert len(eng.aot_engines[0].long_memories_indexes) == FORMER + LATTER, "bank not full after warm-up"
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # The K-step block is timed `--repeat` times back to back (each block bracketed by barrier + synchronize, max over
    # ranks); the reported value is the MEDIAN block, min / max go into config.repeat.  One block of 20 steps is 30 ms.
    l0 = eng.launch_count
    blocks = [timed(frames_dev, args.steps, e2e=False) for _ in range(args.repeat)]
    launches = (eng.launch_count - l0) // args.repeat
    ms = sorted(blocks)[len(blocks) // 2]
    # ---- end-to-end run: pinned host frames in, uint8 label map out, copies inside the timed region ----
    blocks_e2e = [timed(frames_pin, args.steps, e2e=True) for _ in range(args.repeat)]
    ms_e2e = sorted(blocks_e2e)[len(blocks_e2e) // 2]
    # ---- two clips in flight per GPU (additional leg, not the headline): a second engine with its own memory bank on
    # a second stream, both sharing the weights; one host thread issues frame i of clip A, then frame i of clip B.  The
    # per-frame chain of one clip is a sequence of small latency-bound launches, so a second independent chain fills the
    # SMs the first one leaves idle (tools/bench_c5.py --in-flight 2 is the evaluator-level version of the same).
    two = None
    if args.clips_in_flight >= 2:
        two = measure_two_clips(eng, dev, rank, world, args, frames_dev, label0, fill, barrier, build_engine)
    clocks = sampler.stop() if rank == 0 else None

    roof = None
    cpu = None
    if rank == 0:
        roof = measure_attention_roofline(eng, dev, args)
        roof["hbm_kernels"] = measure_hbm_kernels(eng, dev)
        roof["hbm_peak"] = {"value": peaks()["hbm_gbs"], "unit": "GB/s", "source": peaks()["source"]}
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline(sample_frames=args.cpu_frames)
    if rank == 0:
        fps = world * args.steps / (ms / 1e3)
        fps_e2e = world * args.steps / (ms_e2e / 1e3)
        out = {
            "metric": METRIC, "value": round(fps, 3), "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": nwarm, "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "fp16" if _capi.op_dtype() == torch.float16 else "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "parallelism": f"clip-sharded x{world}", "attn_impl": args.attn,
                       "repeat": {"blocks": args.repeat, "steps_per_block": args.steps, "reported": "median block",
                                  "ms_per_step_min": round(min(blocks) / args.steps, 4),
                                  "ms_per_step_max": round(max(blocks) / args.steps, 4),
                                  "e2e_ms_per_step_min": round(min(blocks_e2e) / args.steps, 4),
                                  "e2e_ms_per_step_max": round(max(blocks_e2e) / args.steps, 4)},
                       "bank": "T=8 = FORMER_MEM_LEN 1 + LATTER_MEM_LEN 7 as BASELINE.json names it; the reference's shipped "
                               "eval script uses 1 + 8 (T=9): bench.py --latter 8 runs that setting",
                       "l2": "per-frame working set (banks 3x37 MB + activations + 150 MB attention workspace) "
                             "exceeds the 126 MB L2; no explicit flush",
                       "pipeline": ((f"every {EG}th step = prefetch_n(frames i+{EG} .. i+{2 * EG - 1}: ONE pass of the image encoder over "
                                     f"{EG} images on the engine's side stream); each step = propagate(frame i) + "
                                     "update_memory(frame i); every frame is encoded exactly once, inside the timed region "
                                     "of the block that issues it") if PAIRS else
                                    ("each step = prefetch(frame i+1: image encoder on the engine's side stream) + "
                                     "propagate(frame i) + update_memory(frame i); every frame is encoded exactly once")
                                    if PREFETCH else "no cross-frame prefetch")},
            "e2e": {"value": round(fps_e2e, 3), "unit": "frames/s", "h2d_bytes_per_step": 3 * H * W * 4,
                    "d2h_bytes_per_step": H * W},
            "gpu_launches": int(launches) * world,
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu,
        }
        out["config"]["clips_in_flight_per_gpu"] = 1
        if two is not None:
            out["two_clips_in_flight"] = two
        print(json.dumps(out), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def measure_two_clips(eng, dev, rank, world, args, frames_a, label0, fill, barrier, build_engine):
    """Whole-GPU throughput with TWO independent clips in flight (device-resident frames, same c3 workload per clip).
    Reported next to the single-clip headline, never instead of it."""
    from rmem_b200.synth import synthetic_frames
    ring = frames_a.shape[0] - 1
    NC = args.clips_in_flight
    EG = args.enc_group
    engs, srcs = [eng], [frames_a]
    for k in range(1, NC):
        srcs.append(synthetic_frames(ring + 1, H, W, seed=5000 * k + rank).to(dev))
        engs.append(build_engine("deaotengine", aot_model=eng.AOT, long_term_mem_gap=GAP))
    streams = [torch.cuda.Stream(device=dev) for _ in range(NC)]
    main = torch.cuda.current_stream()

    def step(k, i):
        e, src = engs[k], srcs[k]
        with torch.cuda.stream(streams[k]):
            if PREFETCH and EG > 1:
                if i % EG == 0:
                    e.prefetch_n([src[1 + (i + EG + j) % ring: 2 + (i + EG + j) % ring] for j in range(EG)])
            elif PREFETCH:
                e.prefetch(src[1 + (i + 1) % ring: 2 + (i + 1) % ring])
            lab = e.propagate_label(src[1 + i % ring: 2 + i % ring], output_size=(H, W))
            e.update_memory(lab)

    for k in range(NC):
        streams[k].wait_stream(main)
        with torch.cuda.stream(streams[k]):
            engs[k].restart_engine()
            engs[k].long_term_mem_gap = GAP
            engs[k].add_reference_frame(srcs[k][0:1], label0.int().to(dev), obj_nums=[N_OBJ], frame_step=0)
    nwarm = max(fill, args.warmup)
    nwarm += (-nwarm) % 4
    for i in range(nwarm):
        for k in range(NC):
            step(k, i)
    torch.cuda.synchronize()
    pos = nwarm                                         # the frame index runs on across the timed blocks
    for e in engs:
        assert len(e.aot_engines[0].long_memories_indexes) == FORMER + LATTER, "bank not full after warm-up"
    blocks = []
    for _ in range(args.repeat):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        for k in range(NC):
            streams[k].wait_event(e0)
        for i in range(pos, pos + args.steps):
            for k in range(NC):
                step(k, i)
        pos += args.steps
        for k in range(NC):
            main.wait_stream(streams[k])
        e1.record(main)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        blocks.append(ms)
    ms = sorted(blocks)[len(blocks) // 2]
    print(f"[bench] rank {rank}/{world} {NC} clips in flight: {ms / (NC * args.steps):.4f} ms/frame", file=sys.stderr, flush=True)
    return {"value": round(world * NC * args.steps / (ms / 1e3), 3), "unit": "frames/s", "clips_in_flight_per_gpu": NC,
            "ms_per_frame": round(ms / (NC * args.steps), 4),
            "ms_per_frame_min": round(min(blocks) / (NC * args.steps), 4),
            "ms_per_frame_max": round(max(blocks) / (NC * args.steps), 4),
            "what": f"same c3 workload, {NC} independent clips per GPU (one engine + memory bank + stream each, shared "
                    "weights), frames resident in HBM, all frames of both clips / device time between two events on the "
                    "issuing stream (median block, max over ranks); per-clip latency is what `ms_per_step` reports"}


def measure_attention_roofline(eng, dev, args):
    """Time the dominant kernel -- one GPM layer's long-term attention at c3 (T=8) -- ALONE, on the REAL operands of all
    three layers: Q (= this frame's K), the restricted K / V^T banks and the slot table are taken from the engine's
    state after the timed run (rmem_engine_layer_memory), so peaked layer-1/2 score distributions are what is timed.
    Kernel time = CUDA events recorded by the library immediately around the main kernel launch on its stream; the whole
    op (qprep + seed + kernel + combine) is timed between two events; a 256 MB write flushes L2 between launches."""
    import ctypes as C
    from rmem_b200 import _capi, ops as K
    lib = _capi.load()
    pk = peaks()
    g = torch.Generator().manual_seed(0)
    HW, h, w = HW_TOK, 31, 54
    OP = _capi.op_dtype()
    impl = ATTN_IMPLS[args.attn]
    scale = 1.0 / math.sqrt(128)
    sub = eng.aot_engines[0]
    mems = [sub.layer_memory(l) for l in range(3)]
    T, nslots, HWp = len(mems[0]["slots"]), mems[0]["nslots"], mems[0]["HWp"]
    assert T == FORMER + LATTER, "bank not full"
    flops = 2.0 * HW * (T * HW) * (128 + 1024)
    pe_cur = torch.randn(128, generator=g).to(dev) * 0.05          # same magnitude as cur_pos_emb / mem_pos_emb
    pe_mem = torch.randn(4, 128, generator=g).to(dev) * 0.05
    gate = torch.randn(HW, 1024, generator=g).to(dev).to(OP)
    qt = torch.empty(HW, 128, dtype=OP, device=dev)
    qbias = torch.zeros(HW, T, dtype=torch.float32, device=dev)
    out = torch.empty(HW, 1024, dtype=OP, device=dev)
    mass = torch.empty(HW, T, dtype=torch.float32, device=dev)
    nbytes = C.c_size_t()
    _capi.check(lib.rmem_long_attn_workspace_bytes(impl, HW, HWp, nslots, 1024, C.byref(nbytes)))
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    pes = (C.c_int * T)(*K.temporal_pe_slots(T, 4))
    st = _capi.stream_ptr()

    def launch(l):
        m = mems[l]
        sl = (C.c_int * T)(*m["slots"])
        _capi.check(lib.rmem_qprep_fwd(_capi.ptr(m["q_last"]), C.c_longlong(128), _capi.ptr(pe_cur), _capi.ptr(pe_mem),
                                       pes, T, C.c_float(scale), _capi.ptr(qt), _capi.ptr(qbias), HW, 128, st))
        _capi.check(lib.rmem_long_attn_grid_fwd(impl, _capi.ptr(qt), _capi.ptr(qbias), _capi.ptr(m["kbank"]),
                                                _capi.ptr(m["vtbank"]), nslots, T, sl, HW, HWp, 128, 1024,
                                                C.c_float(scale), _capi.ptr(gate), C.c_longlong(1024), _capi.ptr(out),
                                                C.c_longlong(1024), _capi.ptr(mass), h, w, _capi.ptr(ws),
                                                C.c_size_t(nbytes.value), st))

    for l in range(3):
        launch(l)
    torch.cuda.synchronize()
    iters = 12                                                       # per layer
    per_layer_k, per_layer_op, fallback = [], [], []
    has_events = impl in (ATTN_IMPLS["tc2"], ATTN_IMPLS["tc3"], ATTN_IMPLS["tc4"])
    for l in range(3):
        ks, ops = [], []
        for _ in range(iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); b.record()                                   # create the underlying cudaEvent_t handles
            if has_events:
                _capi.check(lib.rmem_debug_attn_events(C.c_void_p(a.cuda_event), C.c_void_p(b.cuda_event)))
            o0.record()
            launch(l)
            o1.record()
            if has_events:
                _capi.check(lib.rmem_debug_attn_events(None, None))
            torch.cuda.synchronize()
            ops.append(o0.elapsed_time(o1))
            if has_events:
                ks.append(a.elapsed_time(b))
        per_layer_op.append(sorted(ops)[len(ops) // 2])
        per_layer_k.append(sorted(ks)[len(ks) // 2] if ks else per_layer_op[-1])
        if args.attn == "tc4":                                       # did these operands leave the column kernel's fp16 range?
            fallback.append(int(ws[:4].view(torch.int32).item()))
    ms = sum(per_layer_k) / 3
    ms_op = sum(per_layer_op) / 3
    ach = flops / (ms * 1e-3) / 1e12
    kname = {"tc3": "long_attn_tc3_kernel", "tc2": "long_attn_tc2_kernel",
             "tc4": "long_attn_tc4_kernel + guarded long_attn_tc3_kernel fallback"}.get(args.attn, "long_term_attention op")
    return {"bound": "tensor", "kernel": f"{kname} (c3 GPM layer, T={T}, real operands of layers 0/1/2)",
            "achieved": round(ach, 2), "peak": pk["tf_burst"], "unit": "TFLOP/s", "frac": round(ach / pk["tf_burst"], 4),
            "peak_source": pk["source"] + " burst (kernel timed alone)", "ms_per_launch": round(ms, 4),
            "ms_per_launch_by_layer": [round(x, 4) for x in per_layer_k],
            "frac_min": round(flops / (max(per_layer_k) * 1e-3) / 1e12 / pk["tf_burst"], 4),
            "frac_max": round(flops / (min(per_layer_k) * 1e-3) / 1e12 / pk["tf_burst"], 4),
            "op_ms_per_launch": round(ms_op, 4), "op_ms_per_launch_by_layer": [round(x, 4) for x in per_layer_op],
            "op_frac": round(flops / (ms_op * 1e-3) / 1e12 / pk["tf_burst"], 4),
            **({"fallback_taken_by_layer": fallback} if fallback else {}),
            "timing": f"median of {iters} launches per layer on the engine's own steady-state Q / K-bank / V-bank of each "
                      "GPM layer (frac = mean over the three layers); kernel = CUDA events recorded by the library "
                      "immediately around the main kernel launch on its stream; op = qprep + seed + kernel + combine "
                      "between two events; 256 MB L2 flush between launches",
            "algorithmic_flops_per_launch": flops, "traffic": ncu_traffic()}


def measure_hbm_kernels(eng, dev):
    """HBM-bound kernels of the path, each timed ALONE at its c3 shape through the op-level C ABI (median of 9 launches,
    CUDA events around the single launch on the launch stream, 256 MB L2 flush before each: cold-cache, and the ~2 us
    launch latency is inside the time -- a floor for the fraction, not a ceiling).  achieved = ALGORITHMIC bytes (inputs
    read once + outputs written once, as listed) / time; peak = MEASURED_PEAKS.json hbm_gbs."""
    from rmem_b200 import _capi, ops as K
    from rmem_b200.synth import synthetic_frames
    from rmem_b200.weights import pack_model
    pk = peaks()
    OP = _capi.op_dtype()
    g = torch.Generator().manual_seed(1)
    h, w, HW = 31, 54, HW_TOK
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = []

    def timed(name, bytes_, fn, what):
        fn()
        ts = []
        for _ in range(9):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        us = sorted(ts)[len(ts) // 2]
        gbs = bytes_ / us / 1e3
        out.append({"kernel": name, "algorithmic_bytes": int(bytes_), "bytes_are": what, "us": round(us, 2),
                    "achieved_gbs": round(gbs, 1), "frac": round(gbs / pk["hbm_gbs"], 4)})

    x = torch.randn(HW, 1024, generator=g).to(dev).to(OP)
    dw = torch.randn(25, 1024, generator=g).to(dev)
    timed("dwconv5_kernel", 2 * HW * 1024 * 2, lambda: K.dwconv5x5(x, dw, h, w), "gated [HW,1024] fp16 in + out")
    xf = torch.randn(HW, 256, generator=g).to(dev)
    gm, bt = torch.ones(256, device=dev), torch.zeros(256, device=dev)
    timed("layernorm_kernel", HW * 256 * (4 + 2), lambda: K.layernorm(xf, gm, bt), "[HW,256] fp32 in, fp16 out")
    lg = [torch.randn(11, 121, 213, generator=g).to(dev)]
    timed("mask_head_kernel", 11 * 121 * 213 * 4 + H * W, lambda: K.mask_head(lg, H, W, want_logits=False),
          "1/4-res logits fp32 in, uint8 label map out")
    lab = eng.propagate_label(synthetic_frames(2, H, W, seed=3)[1:2].to(dev), output_size=(H, W))[0, 0].contiguous()
    pw = pack_model(eng.AOT.weights_state if hasattr(eng.AOT, "weights_state") else _bench_sd(), "r50_deaotl")
    wts = {k: pw[k].to(dev) for k in ("idbank.w", "idbank.b", "idbank.prefix", "idbank.prefix_rows", "id_norm.g", "id_norm.b",
                                      "enc.conv1.w", "enc.conv1.b")}
    timed("idbank_kernel", H * W + HW * 256 * 4 + pw["idbank.prefix_rows"].numel() * 4,
          lambda: K.id_embedding(lab, wts["idbank.w"], wts["idbank.b"], wts["id_norm.g"], wts["id_norm.b"], True,
                                 prefix=wts["idbank.prefix"], prefix_rows=wts["idbank.prefix_rows"]),
          "uint8 label map + row-prefix table (3.8 MB, each entry at most once) in, [HW,256] fp32 out; steady-state "
          "network labels")
    img = synthetic_frames(1, H, W, seed=5).to(dev)
    H1, W1 = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    timed("gemm_tc_kernel<64> (conv1 stem) + pack_image_padded_kernel", 3 * H * W * 4 + 2 * (H + 6) * (W + 8) * 8 * 2 + H1 * W1 * 64 * 2,
          lambda: K.stem_conv(img, wts["enc.conv1.w"], wts["enc.conv1.b"]),
          "fp32 frame in, padded NHWC8 fp16 written + read once, [P1,64] fp16 out (two launches)")
    return out


def _bench_sd():
    from rmem_b200.synth import make_state_dict
    return make_state_dict("r50_deaotl", seed=0, sharpen=4.0)


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the attention kernel, from the committed
    `ncu --set full` capture (profiles/attn_traffic.json, written by tools/ncu_traffic.py); None if absent."""
    p = os.path.join(ROOT, "profiles", "attn_traffic.json")
    if not os.path.exists(p):
        return None
    try:
        d = json.load(open(p))
        return {"bytes_per_launch": d["dram_bytes_per_launch"], "source": d["source"]}
    except Exception:
        return None


def cpu_baseline(sample_frames: int):
    """CPU oracle port of the reference path (oracle/rmem_oracle.py) on this box's host cores, bounded sample."""
    from oracle import rmem_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = O.make_state_dict("r50_deaotl", seed=0, sharpen=4.0)
    frames = O.synthetic_frames(4, H, W, seed=1000)
    label0 = O.synthetic_label(H, W, N_OBJ)
    # fill the bank quickly with gap=1 so the timed frames see T=8, then switch to the workload's gap
    eng = O.OracleEngine(sd, O.OracleConfig(former_mem_len=FORMER, latter_mem_len=LATTER), long_term_mem_gap=1)
    with torch.no_grad():
        eng.add_reference_frame(frames[0:1], label0, obj_nums=[N_OBJ], frame_step=0)
        for i in range(FORMER + LATTER):
            lg = eng.match_propogate_one_frame(frames[1 + i % 3: 2 + i % 3], output_size=(H, W))
            eng.update_memory(O.logits_to_label(lg))
        eng.long_term_mem_gap = GAP
        t0 = time.perf_counter()
        for i in range(sample_frames):
            lg = eng.match_propogate_one_frame(frames[1 + i % 3: 2 + i % 3], output_size=(H, W))
            eng.update_memory(O.logits_to_label(lg))
        dt = time.perf_counter() - t0
    return {"value": round(sample_frames / dt, 4), "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"{sample_frames} propagated frames of the same c3 workload (bank full, T=8), fp32 torch CPU"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    # one "step" = a bounded sample of one propagated frame; K steps timed after W warm-up frames
    from oracle import rmem_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = O.make_state_dict("r50_deaotl", seed=0, sharpen=4.0)
    frames = O.synthetic_frames(4, H, W, seed=1000)
    label0 = O.synthetic_label(H, W, N_OBJ)
    eng = O.OracleEngine(sd, O.OracleConfig(former_mem_len=FORMER, latter_mem_len=LATTER), long_term_mem_gap=1)
    steps = min(args.steps, args.ref_max_steps)
    warm = min(max(args.warmup, 3), 3)
    with torch.no_grad():
        eng.add_reference_frame(frames[0:1], label0, obj_nums=[N_OBJ], frame_step=0)
        for i in range(FORMER + LATTER):
            lg = eng.match_propogate_one_frame(frames[1 + i % 3: 2 + i % 3], output_size=(H, W))
            eng.update_memory(O.logits_to_label(lg))
        eng.long_term_mem_gap = GAP
        for i in range(warm):
            lg = eng.match_propogate_one_frame(frames[1 + i % 3: 2 + i % 3], output_size=(H, W))
            eng.update_memory(O.logits_to_label(lg))
        t0 = time.perf_counter()
        for i in range(steps):
            lg = eng.match_propogate_one_frame(frames[1 + i % 3: 2 + i % 3], output_size=(H, W))
            eng.update_memory(O.logits_to_label(lg))
        dt = time.perf_counter() - t0
    fps = steps / dt
    out = {"impl": "reference", "metric": METRIC, "value": round(fps, 4), "unit": "frames/s", "n_gpus": world,
           "steps": steps, "warmup": warm, "ms_per_step": round(dt / steps * 1e3, 2), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOAD, "parallelism": "cpu, rank 0 only"},
           "cpu_baseline": {"value": round(fps, 4), "unit": "frames/s", "cores": cores, "kind": "port",
                            "sample": f"{steps} propagated frames (bank full, T=8), fp32 torch CPU, all host threads"},
           "e2e": {"value": round(fps, 4), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    global LATTER, WORKLOAD
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--attn", default=os.environ.get("RMEM_ATTN", "tc4"), choices=["tc4", "tc3", "tc2", "dense"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-frames", type=int, default=10)
    ap.add_argument("--ref-max-steps", type=int, default=20)
    ap.add_argument("--repeat", type=int, default=10, help="timed blocks of --steps steps; the median block is reported")
    ap.add_argument("--enc-group", type=int, default=int(os.environ.get("RMEM_BENCH_ENC_GROUP", "2")), choices=[1, 2, 4],
                    help="frames per pass of the image encoder (rmem_engine_prefetch_n, issued that many frames ahead); "
                         "1 = single-frame prefetch")
    ap.add_argument("--clips-in-flight", type=int, default=2,
                    help="2 = also time two independent clips per GPU (extra key two_clips_in_flight); 1 = skip that leg")
    ap.add_argument("--latter", type=int, default=LATTER, help="LATTER_MEM_LEN (7 = T=8 as BASELINE.json names it)")
    args = ap.parse_args()
    if args.latter != LATTER:
        LATTER = args.latter
        WORKLOAD = WORKLOAD.replace("T=8 (former 1 + latter 7)", f"T={1 + LATTER} (former 1 + latter {LATTER})")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
