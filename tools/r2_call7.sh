#!/bin/bash
# Round 2, GPU call 7: per-warp hand-over slots + gather-form stage 2a.
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/gputests7.log 2>&1
echo "gpu tests exit code $?" | tee -a gpurun_out/gputests7.log
timeout 120 python tools/ggs_stage_probe.py 20 2048 paired > gpurun_out/probe7_cfg3_paired.txt 2>&1
timeout 300 python tools/ggs_stage_probe.py 80 4096 paired > gpurun_out/probe7_cfg5_paired.txt 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --ggs-layout paired > gpurun_out/bench7_cfg3_paired.json 2> gpurun_out/bench7_cfg3_paired.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --seqs-per-gpu 8 --ggs-layout paired > gpurun_out/bench7_b8_paired.json 2> gpurun_out/bench7_b8_paired.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload cfg5 --ggs-layout paired > gpurun_out/bench7_cfg5_paired.json 2> gpurun_out/bench7_cfg5_paired.err
tail -n 4 gpurun_out/gputests7.log
cat gpurun_out/probe7_cfg3_paired.txt gpurun_out/probe7_cfg5_paired.txt
for f in gpurun_out/bench7_*.json; do echo "$f: $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d.get('kernel_ms_per_loop'), (d.get('roofline') or {}).get('frac'))" 2>&1)"; done
