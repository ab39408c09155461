#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list, ncu full capture of the attention kernel.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
timeout 600 python bench.py --steps 100 --warmup 10 --attn dense --no-cpu-baseline > gpurun_out/bench_dense.json 2> gpurun_out/bench_dense.err
cat gpurun_out/bench_dense.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python tools/profile_frame.py --frames 5 > gpurun_out/launches.log 2>&1
tail -2 gpurun_out/launches.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:long_attn_tc_kernel -c 2 -o gpurun_out/attn_tc python tools/profile_frame.py --frames 1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out
