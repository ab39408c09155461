#!/usr/bin/env python
"""Host-side issue time of one steady-state frame (propagate + update) with an empty launch queue, vs its GPU time.
If the two are close the path is launch-bound on this host and multi-process scaling will suffer."""
import os, sys, time, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import bench as B
from rmem_b200.engine import DeAOTModel, RmemConfig, build_engine
from rmem_b200.synth import make_state_dict, synthetic_frames, synthetic_label

dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
sd = make_state_dict("r50_deaotl", seed=0, sharpen=4.0)
cfg = RmemConfig(former_mem_len=B.FORMER, latter_mem_len=B.LATTER, attn_impl=B.ATTN_IMPLS["tc3"], max_engines=1)
eng = build_engine("deaotengine", aot_model=DeAOTModel(sd, cfg, dev), long_term_mem_gap=B.GAP)
frames = synthetic_frames(9, B.H, B.W, seed=1000).to(dev); label0 = synthetic_label(B.H, B.W, B.N_OBJ)
eng.restart_engine(); eng.long_term_mem_gap = B.GAP
eng.add_reference_frame(frames[0:1], label0.int().to(dev), obj_nums=[B.N_OBJ], frame_step=0)
def step(i):
    lab = eng.propagate_label(frames[1 + i % 8: 2 + i % 8], output_size=(B.H, B.W)); eng.update_memory(lab)
for i in range(60): step(i)
torch.cuda.synchronize()
cpu, gpu = [], []
for i in range(40):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(); step(60 + i); e1.record(); t1 = time.perf_counter()
    torch.cuda.synchronize()
    cpu.append((t1 - t0) * 1e3); gpu.append(e0.elapsed_time(e1))
print(f"host issue time per frame: median {statistics.median(cpu):.3f} ms (min {min(cpu):.3f})   GPU time per isolated frame: median {statistics.median(gpu):.3f} ms")
print("cores:", os.cpu_count(), "affinity:", len(os.sched_getaffinity(0)))
