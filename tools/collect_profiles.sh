#!/bin/bash
# Copy the judged artefacts of the last tools/gpu_round.sh visit from gpurun_out/ (scratch) into profiles/ (tracked).
set -e
R=${1:-r01}
cp gpurun_out/bench.json profiles/${R}_bench_final.json
cp gpurun_out/bench_ref.json profiles/${R}_bench_reference_arm.json
cp gpurun_out/stage_times.txt profiles/${R}_stage_times.txt
cp gpurun_out/launches.csv profiles/${R}_launches_c3_final.csv
python tools/summarize_launches.py gpurun_out/launches.csv 5 > profiles/${R}_launches_c3_final_summary.txt
[ -f gpurun_out/bench_2gpu.json ] && grep '^{' gpurun_out/bench_2gpu.json > profiles/${R}_bench_2gpu.json
ncu -i gpurun_out/attn.ncu-rep --page details --csv > profiles/${R}_attn_final_ncu_details.csv
python tools/ncu_traffic.py gpurun_out/attn.ncu-rep long_attn_tc2_kernel profiles/attn_traffic.json
python - <<PY
import csv, io, subprocess
txt = subprocess.run(["ncu", "-i", "gpurun_out/attn.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active"]
with open("profiles/${R}_attn_final_ncu_key_metrics.txt", "w") as f:
    f.write("ncu --set full --clock-control none, tools/profile_frame.py --frames 1 (c3 steady state, bank full); cold-cache, serialised launches\n")
    for r in rows[2:]:
        f.write("\n")
        for k in keys:
            if k in hdr:
                i = hdr.index(k)
                f.write(f"{k:80s} {r[i]} {units[i]}\n")
print(open("profiles/${R}_attn_final_ncu_key_metrics.txt").read()[:3000])
PY
