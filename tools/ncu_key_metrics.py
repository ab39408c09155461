#!/usr/bin/env python
"""Key metrics of every launch in an .ncu-rep (ncu --set full) as a small text table: duration, tensor-pipe activity,
SM / L2 / DRAM throughput, DRAM bytes, registers, shared memory, grid.   python tools/ncu_key_metrics.py rep.ncu-rep out.txt"""
import csv, io, subprocess, sys

rep, out = sys.argv[1:3]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "launch__cluster_size", "launch__waves_per_multiprocessor"]
ik = hdr.index("Kernel Name")
with open(out, "w") as f:
    f.write(f"ncu --set full --clock-control none ({rep.split('/')[-1]}); one steady-state c3 frame (tools/profile_frame.py --frames 1); "
            "cold-cache, serialised launches: compare shares and pipe percentages, not absolute times\n")
    for r in rows[2:]:
        f.write(f"\n{r[ik][:110]}\n")
        for k in keys:
            if k in hdr:
                i = hdr.index(k)
                f.write(f"    {k:72s} {r[i]} {units[i]}\n")
print(open(out).read()[:2500])
