#!/bin/bash
# AOT path on the fused tcgen05 multi-head attention: op parity, engine goldens, c2 full size, frames/s of c2/c3/c4
mkdir -p gpurun_out
timeout 300 python tests/mha_tc_check.py > gpurun_out/mha_check.log 2>&1; echo "mha check rc=$?"
cat gpurun_out/mha_check.log | tail -12
timeout 900 python -m pytest tests -m gpu -x -q -k "aot or mha or c2 or evaluator or long_clip" > gpurun_out/pytest_aot.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_aot.log
timeout 300 python tools/fps_configs.py --frames 60 > gpurun_out/fps_configs.txt 2>&1
cat gpurun_out/fps_configs.txt | tail -4
