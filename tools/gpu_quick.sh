#!/bin/bash
# quick loop: gemm + attention parity, bench (no cpu baseline), warm launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_tc_gpu.py tests/test_ops_gpu.py -x -q > gpurun_out/pytest_quick.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_quick.log
tail -4 gpurun_out/pytest_quick.log
timeout 300 python bench.py --steps 60 --warmup 10 --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/bench_q.json'));print('fps',d['value'],'e2e',d['e2e']['value'],'attn ms',d['roofline']['ms_per_launch'],'frac',d['roofline']['frac'],'launches/step',d['gpu_launches']/d['steps'])"; tail -3 gpurun_out/bench_q.err
timeout 600 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_warm.csv python tools/profile_frame.py --frames 5 > gpurun_out/launches_warm.log 2>&1
tail -1 gpurun_out/launches_warm.log
