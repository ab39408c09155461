#!/bin/bash
# column attention kernel (RMEM_ATTN_TC4): op parity (incl. the guarded fallback), phase trace, stage times, A/B bench against tc3
mkdir -p gpurun_out
RMEM_ATTN_IMPL=4 RMEM_ATTN_SEED=1 timeout 300 python tests/tc_attn_check.py > gpurun_out/tc4_check.log 2>&1; echo "tc4 check rc=$?"
cut -c1-200 gpurun_out/tc4_check.log | tail -8
timeout 200 python tools/trace_attn4.py 2>&1 | tail -12
timeout 300 python tools/profile_frame.py --frames 40 --stages --attn tc4 2>&1 | grep -E "gpm.long.attn|total"
timeout 600 python tools/bench_brief.py tc3 --attn tc3
timeout 600 python tools/bench_brief.py tc4 --attn tc4
