#!/bin/bash
# iteration loop: attention parity + op/gemm parity + bench + warm (no cache flush) launch list
mkdir -p gpurun_out
RMEM_ATTN_IMPL=2 timeout 600 python tests/tc_attn_check.py > gpurun_out/attn2_check.log 2>&1; echo "rc=$?" >> gpurun_out/attn2_check.log
cut -c1-150 gpurun_out/attn2_check.log | tail -9
timeout 900 python -m pytest tests/test_gemm_tc_gpu.py tests/test_ops_gpu.py tests/test_engine_gpu.py -x -q > gpurun_out/pytest_quick.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_quick.log
tail -4 gpurun_out/pytest_quick.log
timeout 300 python bench.py --steps 60 --warmup 10 --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/bench_q.json'));r=d['roofline'];print('fps',d['value'],'e2e',d['e2e']['value'],'kernel ms',r['ms_per_launch'],r['frac'],'op',r['op_ms_per_launch'],r['op_frac'])"; tail -3 gpurun_out/bench_q.err
timeout 600 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_warm.csv python tools/profile_frame.py --frames 5 > gpurun_out/launches_warm.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_warm.csv 2>/dev/null | head -30
