#!/bin/bash
# c5 driver on 8 GPUs: one clip and two clips in flight per GPU
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/bench_c5.py --clips 16 --frames 2000 > gpurun_out/r02_c5_8gpu.json 2> gpurun_out/r02_c5_8gpu.err; echo "c5 rc=$?"
tail -1 gpurun_out/r02_c5_8gpu.json | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tools/bench_c5.py --clips 32 --frames 2000 --in-flight 2 > gpurun_out/r02_c5_8gpu_inflight2.json 2> gpurun_out/r02_c5_8gpu_inflight2.err; echo "c5 x2 rc=$?"
tail -1 gpurun_out/r02_c5_8gpu_inflight2.json | cut -c1-300
