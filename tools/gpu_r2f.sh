#!/bin/bash
mkdir -p gpurun_out
python tools/bench_brief.py default --clips-in-flight 1
RMEM_SIDE_PDL=1 python tools/bench_brief.py side_pdl --clips-in-flight 1
RMEM_ENC_PRIO=1 python tools/bench_brief.py enc_prio --clips-in-flight 1
RMEM_ENC_PRIO=1 RMEM_SIDE_PDL=1 python tools/bench_brief.py enc_prio_side_pdl --clips-in-flight 1
RMEM_BRANCH_PAR=0 python tools/bench_brief.py no_branch_par --clips-in-flight 1
