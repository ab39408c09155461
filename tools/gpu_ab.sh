#!/bin/bash
# parity (all GPU tests) + A/B of the engine switches through bench.py (20-step blocks x 10, median)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python tools/bench_brief.py default
RMEM_DW_TW=9 python tools/bench_brief.py dw_tw9
RMEM_DW_TW=6 python tools/bench_brief.py dw_tw6
RMEM_DW_TW=27 python tools/bench_brief.py dw_tw27
python tools/bench_brief.py default_again
timeout 300 python tools/fps_configs.py --frames 60 2>&1 | tail -3
