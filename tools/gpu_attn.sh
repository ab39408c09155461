#!/bin/bash
mkdir -p gpurun_out
RMEM_ATTN_IMPL=2 timeout 600 python tests/tc_attn_check.py > gpurun_out/attn2_check.log 2>&1; echo "rc=$?" >> gpurun_out/attn2_check.log
cut -c1-170 gpurun_out/attn2_check.log | tail -9
RMEM_TRACE_CTA=95 python tools/trace_attn.py > gpurun_out/attn_trace95.txt 2>&1; tail -14 gpurun_out/attn_trace95.txt
timeout 300 python bench.py --steps 60 --warmup 10 --no-cpu-baseline > gpurun_out/bench_tc2.json 2> gpurun_out/bench_tc2.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/bench_tc2.json'));r=d['roofline'];print('fps',d['value'],'e2e',d['e2e']['value'],'kernel ms',r['ms_per_launch'],r['frac'],'op',r['op_ms_per_launch'],r['op_frac'])"; tail -5 gpurun_out/bench_tc2.err
