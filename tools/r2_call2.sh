#!/bin/bash
# Round 2, GPU call 2: the rewritten GGS iteration tail (exchange through flag-carrying words): all GPU tests, the extended
# exchange microbenchmark, per-stage probe and bench at config 3 for both layouts, group-size / CTA-count sweep.
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/gputests.log 2>&1
echo "gpu tests exit code $?" | tee -a gpurun_out/gputests.log
timeout 300 build/xchg_probe 148 20000 > gpurun_out/xchg_probe2.txt 2>&1
for layout in plain paired; do
  timeout 120 python tools/ggs_stage_probe.py 20 2048 $layout > gpurun_out/probe_cfg3_$layout.txt 2>&1
done
for g in 6 8 10 16 24; do
  PDB_GGS_GROUP=$g timeout 120 python tools/ggs_stage_probe.py 20 2048 paired > gpurun_out/probe_cfg3_paired_g$g.txt 2>&1
done
for c in 37 74 111; do
  PDB_GGS_CPP=$c timeout 120 python tools/ggs_stage_probe.py 20 2048 paired > gpurun_out/probe_cfg3_paired_c$c.txt 2>&1
done
for layout in plain paired; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --ggs-layout $layout > gpurun_out/bench_cfg3_$layout.json 2> gpurun_out/bench_cfg3_$layout.err
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --seqs-per-gpu 8 --ggs-layout $layout > gpurun_out/bench_b8_$layout.json 2> gpurun_out/bench_b8_$layout.err
done
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload cfg5 --ggs-layout paired > gpurun_out/bench_cfg5_paired.json 2> gpurun_out/bench_cfg5_paired.err
tail -n 12 gpurun_out/gputests.log
cat gpurun_out/xchg_probe2.txt
head -5 gpurun_out/probe_cfg3_*.txt
for f in gpurun_out/bench_*_*.json; do echo "$f: $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d.get('kernel_ms_per_loop'), (d.get('roofline') or {}).get('frac'))" 2>&1)"; done
