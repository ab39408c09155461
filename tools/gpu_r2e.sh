#!/bin/bash
# parity subset + A/B of the layer re-scheduling and the GEMM heuristics
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_gemm_tc_gpu.py tests/test_layer_goldens_gpu.py -m gpu -x -q 2>&1 | tail -4
python tools/bench_brief.py default --clips-in-flight 1
RMEM_GEMM_DEEP=0 python tools/bench_brief.py no_deep --clips-in-flight 1
RMEM_GEMM_WIDE_MIN=120 python tools/bench_brief.py wide120 --clips-in-flight 1
RMEM_GEMM_SPLITK=0 python tools/bench_brief.py no_splitk --clips-in-flight 1
python tools/bench_brief.py default_again --clips-in-flight 2
timeout 300 python tools/timeline_frame.py --frames 4 --json gpurun_out/timeline_1clip_b.json > gpurun_out/timeline_1clip_b.txt 2>&1; echo "timeline rc=$?"
grep "^# " gpurun_out/timeline_1clip_b.txt | head -12
