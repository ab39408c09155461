#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_tc_gpu.py tests/test_engine_gpu.py -m gpu -x -q 2>&1 | tail -15
python tools/bench_brief.py pairs --clips-in-flight 1 --enc-pairs 1
python tools/bench_brief.py single --clips-in-flight 1 --enc-pairs 0
RMEM_SIDE_PDL=1 python tools/bench_brief.py pairs_side_pdl --clips-in-flight 1 --enc-pairs 1
python tools/bench_brief.py pairs_again --clips-in-flight 1 --enc-pairs 1
