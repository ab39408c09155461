#!/bin/bash
# Last GPU visit of round 2 (a few minutes of budget): the training-side slice first, then as much of the full GPU
# suite as fits.  Everything is written to gpurun_out/ as it goes, so a cut-off still leaves the finished parts.
mkdir -p gpurun_out
BUDGET=${BUDGET:-250}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/r4_gpu.txt 2>&1
timeout 120 python -u -m pytest tests/test_training_gpu.py -m gpu -v -s -x > gpurun_out/r4_train_pytest.log 2>&1
echo "train pytest rc=$? at ${SECONDS}s"; tail -3 gpurun_out/r4_train_pytest.log
timeout 40 python tools/bench_train_loss.py > gpurun_out/r4_train_loss_bench.json 2> gpurun_out/r4_train_loss_bench.err
echo "loss bench rc=$? at ${SECONDS}s"; cat gpurun_out/r4_train_loss_bench.json
LEFT=$((BUDGET - SECONDS))
if [ "$LEFT" -gt 20 ]; then
  timeout $LEFT python -u -m pytest tests -m gpu -x -v --ignore tests/test_training_gpu.py > gpurun_out/r4_pytest_gpu.log 2>&1
  echo "full pytest rc=$? (124 = out of time) at ${SECONDS}s"; tail -2 gpurun_out/r4_pytest_gpu.log
  grep -c PASSED gpurun_out/r4_pytest_gpu.log
fi
