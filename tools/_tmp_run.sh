timeout 600 python -m pytest tests/test_evaluator.py tests/test_engine_gpu.py -m gpu -x -q 2>&1 | tail -3
python tools/bench_brief.py e2e_copy_ahead --clips-in-flight 1
timeout 600 python tools/bench_c5.py --clips 2 --frames 2000 2> /dev/null | tail -1 | cut -c1-330
