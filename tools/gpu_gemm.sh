#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_tc_gpu.py -x -q > gpurun_out/pytest_gemm.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gemm.log
tail -30 gpurun_out/pytest_gemm.log
timeout 300 python tools/bench_gemm.py > gpurun_out/bench_gemm.log 2>&1
cat gpurun_out/bench_gemm.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
