#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_tc_gpu.py tests/test_engine_gpu.py tests/test_ops_gpu.py tests/test_layer_goldens_gpu.py -m gpu -x -q 2>&1 | tail -12
python tools/bench_brief.py gn_fused --clips-in-flight 1
RMEM_GN_FUSED=0 python tools/bench_brief.py gn_separate --clips-in-flight 1
python tools/bench_brief.py gn_fused_again --clips-in-flight 2
