#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/profile_frame.py --frames 40 --stages --attn tc3 2>&1 | grep -E "gpm.long|gpm.self.attn|total" 
timeout 300 python tools/profile_frame.py --frames 40 --stages --attn tc4 2>&1 | grep -E "gpm.long|gpm.self.attn|total"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_tc4.csv python tools/profile_frame.py --frames 3 --attn tc4 > gpurun_out/launches_tc4.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_tc4.csv 3 | grep -E "attn|combine|seed|launches"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --attn tc4 > gpurun_out/bench_tc4.json 2>gpurun_out/bench_tc4.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_tc4.json').read().strip().splitlines()[-1]); r=d['roofline']; print({k:r[k] for k in r if k not in ('timing','hbm_kernels','traffic')})"
