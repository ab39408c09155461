#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -x -q 2>&1 | tail -8
python tools/bench_brief.py group4 --clips-in-flight 1 --enc-group 4
python tools/bench_brief.py group2 --clips-in-flight 1 --enc-group 2
python tools/bench_brief.py group1 --clips-in-flight 1 --enc-group 1
python tools/bench_brief.py group4_again --clips-in-flight 2 --enc-group 4
