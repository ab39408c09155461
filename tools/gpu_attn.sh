#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python bench.py --steps 60 --warmup 10 --no-cpu-baseline --attn tc2 > gpurun_out/bench_tc2.json 2> gpurun_out/bench_tc2.err; echo "bench rc=$?"
cat gpurun_out/bench_tc2.json; tail -5 gpurun_out/bench_tc2.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches3.csv python tools/profile_frame.py --frames 5 > gpurun_out/launches3.log 2>&1
tail -2 gpurun_out/launches3.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:long_attn_tc2_kernel -c 2 -o gpurun_out/attn_tc2 python tools/profile_frame.py --frames 1 > gpurun_out/ncu_full2.log 2>&1
tail -2 gpurun_out/ncu_full2.log
