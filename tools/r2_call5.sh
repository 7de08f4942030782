#!/bin/bash
# Round 2, GPU call 5: stage 3 / 2b restructure, touched-only one-hop exchange, denoiser second thread mapping.
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/gputests5.log 2>&1
echo "gpu tests exit code $?" | tee -a gpurun_out/gputests5.log
timeout 120 python tools/ggs_stage_probe.py 20 2048 paired > gpurun_out/probe5_cfg3_paired.txt 2>&1
timeout 120 python tools/ggs_stage_probe.py 20 2048 plain > gpurun_out/probe5_cfg3_plain.txt 2>&1
timeout 300 python tools/ggs_stage_probe.py 80 4096 paired > gpurun_out/probe5_cfg5_paired.txt 2>&1
timeout 300 python tools/den_stage_probe.py 20 1 > gpurun_out/den_probe5_n20.txt 2>&1
timeout 300 python tools/den_stage_probe.py 80 1 > gpurun_out/den_probe5_n80.txt 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --ggs-layout paired > gpurun_out/bench5_cfg3_paired.json 2> gpurun_out/bench5_cfg3_paired.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --seqs-per-gpu 8 --ggs-layout paired > gpurun_out/bench5_b8_paired.json 2> gpurun_out/bench5_b8_paired.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload cfg5 --ggs-layout paired > gpurun_out/bench5_cfg5_paired.json 2> gpurun_out/bench5_cfg5_paired.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload cfg2 > gpurun_out/bench5_cfg2.json 2> gpurun_out/bench5_cfg2.err
tail -n 6 gpurun_out/gputests5.log
head -8 gpurun_out/probe5_*.txt
cat gpurun_out/den_probe5_n20.txt gpurun_out/den_probe5_n80.txt
for f in gpurun_out/bench5_*.json; do echo "$f: $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d.get('kernel_ms_per_loop'), (d.get('roofline') or {}).get('frac'))" 2>&1)"; done
