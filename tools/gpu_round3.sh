#!/bin/bash
# Evidence visit without the 14-minute ncu --set full pass: parity tests, smoke, bench (both arms), stage times, ncu launch
# list, c5 driver on one GPU, c2 / c4 frame rates.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -1 gpurun_out/bench.json | cut -c1-1800
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -1 gpurun_out/bench_ref.json | cut -c1-300
timeout 300 python tools/profile_frame.py --frames 40 --stages > gpurun_out/stage_times.txt 2>&1
tail -2 gpurun_out/stage_times.txt
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python tools/profile_frame.py --frames 6 > gpurun_out/launches.log 2>&1
tail -1 gpurun_out/launches.log
timeout 600 python tools/bench_c5.py --clips 2 --frames 2000 > gpurun_out/c5_1gpu.json 2> gpurun_out/c5_1gpu.err; tail -1 gpurun_out/c5_1gpu.json | cut -c1-400
timeout 300 python tools/fps_configs.py --frames 60 2>&1 | tail -3 | tee gpurun_out/fps_configs.txt
