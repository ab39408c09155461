#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_tc_gpu.py tests/test_engine_gpu.py -m gpu -x -q 2>&1 | tail -15
timeout 300 python tools/timeline_frame.py --frames 4 --json gpurun_out/timeline_pairs.json > gpurun_out/timeline_pairs.txt 2>&1; echo "timeline rc=$?"
grep "^# " gpurun_out/timeline_pairs.txt | head -14
