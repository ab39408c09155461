#!/bin/bash
# quick visit: selected parity tests + one short bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_full_size_gpu.py tests/test_engine_gpu.py tests/test_evaluator.py -m gpu -x -q -s 2>&1 | grep -E "^\[|passed|failed|Error|error" | tail -14
python tools/bench_brief.py e2e_d2h_stream --clips-in-flight 1
timeout 600 python tools/bench_c5.py --clips 2 --frames 1000 2> /dev/null | tail -1 | cut -c1-300
