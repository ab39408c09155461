#!/bin/bash
# quick visit: selected parity tests + short bench lines (A/B of an environment switch: bash tools/gpu_check.sh VAR=value)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_tc_gpu.py tests/test_engine_gpu.py tests/test_layer_goldens_gpu.py tests/test_ops_gpu.py -m gpu -x -q 2>&1 | tail -6
python tools/bench_brief.py default --clips-in-flight 1
if [ -n "$1" ]; then env "$1" python tools/bench_brief.py "$1" --clips-in-flight 1; python tools/bench_brief.py default_again --clips-in-flight 1; fi
