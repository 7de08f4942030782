#!/bin/bash
# Round 2, GPU call 4: one-hop exchange in the kernel: all GPU tests, stage probes for both exchange modes, benches.
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/gputests4.log 2>&1
echo "gpu tests exit code $?" | tee -a gpurun_out/gputests4.log
for x in 1 0; do
  PDB_GGS_XCH=$x timeout 120 python tools/ggs_stage_probe.py 20 2048 paired > gpurun_out/probe4_cfg3_paired_x$x.txt 2>&1
done
PDB_GGS_XCH=0 PDB_GGS_GROUP=10 timeout 120 python tools/ggs_stage_probe.py 20 2048 paired > gpurun_out/probe4_cfg3_paired_x0_g10.txt 2>&1
timeout 120 python tools/ggs_stage_probe.py 20 2048 plain > gpurun_out/probe4_cfg3_plain_x1.txt 2>&1
timeout 300 python tools/ggs_stage_probe.py 80 4096 paired > gpurun_out/probe4_cfg5_paired_x1.txt 2>&1
for layout in plain paired; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --ggs-layout $layout > gpurun_out/bench4_cfg3_$layout.json 2> gpurun_out/bench4_cfg3_$layout.err
done
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --seqs-per-gpu 8 --ggs-layout paired > gpurun_out/bench4_b8_paired.json 2> gpurun_out/bench4_b8_paired.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload cfg5 --ggs-layout paired > gpurun_out/bench4_cfg5_paired.json 2> gpurun_out/bench4_cfg5_paired.err
tail -n 8 gpurun_out/gputests4.log
head -8 gpurun_out/probe4_*.txt
for f in gpurun_out/bench4_*.json; do echo "$f: $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d.get('kernel_ms_per_loop'), (d.get('roofline') or {}).get('frac'))" 2>&1)"; done
