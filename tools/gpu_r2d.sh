#!/bin/bash
# GEMM tile sweep over the frame's shapes, CUPTI timeline of production frames (1 and 2 clips), bench with 2 / 3 clips in flight
mkdir -p gpurun_out
timeout 600 python tools/tune_gemm.py > gpurun_out/tune_gemm.jsonl 2> gpurun_out/tune_gemm.err; echo "tune rc=$?"
grep "^#" gpurun_out/tune_gemm.jsonl | head -60
timeout 300 python tools/timeline_frame.py --frames 4 --json gpurun_out/timeline_1clip.json > gpurun_out/timeline_1clip.txt 2>&1; echo "timeline rc=$?"
grep "^# " gpurun_out/timeline_1clip.txt | head -50
timeout 300 python tools/timeline_frame.py --frames 4 --clips 2 > gpurun_out/timeline_2clips.txt 2>&1
grep "^# " gpurun_out/timeline_2clips.txt | head -12
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_2clips.json 2> gpurun_out/bench_2clips.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_2clips.json').read().strip().splitlines()[-1])
print('fps',d['value'],'e2e',d['e2e']['value'],'two',d.get('two_clips_in_flight'))
PY
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --clips-in-flight 3 > gpurun_out/bench_3clips.json 2> gpurun_out/bench_3clips.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_3clips.json').read().strip().splitlines()[-1])
print('fps',d['value'],'three',d.get('two_clips_in_flight'))
PY
RMEM_BENCH_PREFETCH=0 python tools/bench_brief.py no_prefetch --clips-in-flight 1
RMEM_AUX_STREAM=0 python tools/bench_brief.py no_aux --clips-in-flight 1
RMEM_BRANCH_PAR=0 python tools/bench_brief.py no_branch_par --clips-in-flight 1
