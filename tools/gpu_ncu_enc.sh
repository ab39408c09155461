#!/bin/bash
# ncu --set full of the tcgen05 GEMM / conv launches of steady-state frames in pair-encoder mode (one encoder pass over two
# images + two frames of GPM / decoder GEMMs), summarised on the box.
mkdir -p gpurun_out
timeout 1500 ncu --profile-from-start off --set full --clock-control none -k regex:gemm_tc_kernel -c 130 \
    -o /tmp/gemm python tools/profile_frame.py --frames 2 > gpurun_out/ncu_gemm.log 2>&1
tail -2 gpurun_out/ncu_gemm.log
python tools/ncu_key_metrics.py /tmp/gemm.ncu-rep gpurun_out/gemm_pairs_ncu_key_metrics.txt > /dev/null 2>&1
ls -la gpurun_out/gemm_pairs_ncu_key_metrics.txt
