#!/bin/bash
mkdir -p gpurun_out
RMEM_ATTN_IMPL=2 timeout 600 python tests/tc_attn_check.py > gpurun_out/attn2_check.log 2>&1; echo "rc=$?" >> gpurun_out/attn2_check.log
cut -c1-160 gpurun_out/attn2_check.log | tail -9
timeout 300 python bench.py --steps 60 --warmup 10 --no-cpu-baseline --attn tc2 > gpurun_out/bench_tc2.json 2> gpurun_out/bench_tc2.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/bench_tc2.json'));print(d['value'],d['e2e']['value'],d['roofline']['ms_per_launch'],d['roofline']['frac'])"; tail -5 gpurun_out/bench_tc2.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python tools/profile_frame.py --frames 5 > gpurun_out/launches.log 2>&1
tail -1 gpurun_out/launches.log
