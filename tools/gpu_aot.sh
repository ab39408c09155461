#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py -x -q -k "aot_engine" > gpurun_out/pytest_aot.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_aot.log
tail -40 gpurun_out/pytest_aot.log
