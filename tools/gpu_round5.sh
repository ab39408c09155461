#!/bin/bash
# Very last GPU visit of round 2 (~1.5 minutes of budget): the 6-launch loss head -- its GPU tests and its timing.
mkdir -p gpurun_out
timeout 50 python -u -m pytest tests/test_training_gpu.py -m gpu -v -s -x > gpurun_out/r5_train_pytest.log 2>&1
echo "train pytest rc=$? at ${SECONDS}s"; tail -3 gpurun_out/r5_train_pytest.log
timeout 25 python tools/bench_train_loss.py > gpurun_out/r5_train_loss_bench.json 2> gpurun_out/r5_train_loss_bench.err
echo "loss bench rc=$? at ${SECONDS}s"; cat gpurun_out/r5_train_loss_bench.json
