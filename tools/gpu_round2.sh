#!/bin/bash
# Round-2 GPU-box visit: parity tests, smoke, bench (both arms), stage times, ncu launch list, ncu --set full of one
# steady-state frame summarised ON THE BOX (the .ncu-rep stays there: gpurun_out/ pulls back at most 64 MiB).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -1 gpurun_out/bench.json | cut -c1-3000
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -1 gpurun_out/bench_ref.json | cut -c1-600
timeout 300 python tools/profile_frame.py --frames 40 --stages > gpurun_out/stage_times.txt 2>&1
tail -3 gpurun_out/stage_times.txt
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python tools/profile_frame.py --frames 5 > gpurun_out/launches.log 2>&1
tail -1 gpurun_out/launches.log
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -o /tmp/frame python tools/profile_frame.py --frames 1 > gpurun_out/ncu_frame.log 2>&1
tail -1 gpurun_out/ncu_frame.log
python tools/ncu_key_metrics.py /tmp/frame.ncu-rep gpurun_out/frame_ncu_key_metrics.txt > /dev/null 2>&1
python tools/ncu_traffic.py /tmp/frame.ncu-rep long_attn_tc4_kernel gpurun_out/attn_traffic.json > /dev/null 2>&1
ls -la gpurun_out | head -40
