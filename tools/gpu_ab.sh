#!/bin/bash
# parity (all GPU tests) + A/B of the engine switches through bench.py (20-step blocks x 10, median)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python tools/bench_brief.py default
RMEM_FUSED_SEED=0 python tools/bench_brief.py no_fused_seed
RMEM_SELF_SEED=0 python tools/bench_brief.py no_self_seed
python tools/bench_brief.py default_again
