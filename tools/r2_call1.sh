#!/bin/bash
# Round 2, GPU call 1: every GPU test (paired layout un-gated, oracle comparisons at BASELINE sizes), the exchange-latency
# microbenchmark, both stream layouts benched at config 3 / config 5 / 8 sequences per GPU, per-stage cycle probe, and the
# baseline ncu captures (raw CSV exported) of the round-1 kernels.  Everything lands in gpurun_out/.
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/gputests.log 2>&1
echo "gpu tests exit code $?" | tee -a gpurun_out/gputests.log
timeout 300 build/xchg_probe 148 20000 > gpurun_out/xchg_probe.txt 2>&1
for layout in plain paired; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --ggs-layout $layout > gpurun_out/bench_cfg3_$layout.json 2> gpurun_out/bench_cfg3_$layout.err
  timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload cfg5 --ggs-layout $layout > gpurun_out/bench_cfg5_$layout.json 2> gpurun_out/bench_cfg5_$layout.err
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --seqs-per-gpu 8 --ggs-layout $layout > gpurun_out/bench_b8_$layout.json 2> gpurun_out/bench_b8_$layout.err
  timeout 120 python tools/ggs_stage_probe.py 20 2048 $layout > gpurun_out/probe_cfg3_$layout.txt 2>&1
  timeout 300 python tools/ggs_stage_probe.py 80 4096 $layout > gpurun_out/probe_cfg5_$layout.txt 2>&1
done
for layout in plain paired; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:ggs_entry -s 2 -c 1 -f -o gpurun_out/ggs_cfg5_$layout \
    python tools/ggs_stage_probe.py 80 4096 $layout > gpurun_out/ncu_cfg5_$layout.log 2>&1
  ncu -i gpurun_out/ggs_cfg5_$layout.ncu-rep --page raw --csv > gpurun_out/ggs_cfg5_${layout}_raw.csv 2>/dev/null
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ggs_entry -s 2 -c 1 -f -o gpurun_out/ggs_cfg3_plain \
  python tools/ggs_stage_probe.py 20 2048 plain > gpurun_out/ncu_cfg3_plain.log 2>&1
ncu -i gpurun_out/ggs_cfg3_plain.ncu-rep --page raw --csv > gpurun_out/ggs_cfg3_plain_raw.csv 2>/dev/null
tail -n 5 gpurun_out/gputests.log
cat gpurun_out/xchg_probe.txt
for f in gpurun_out/bench_*_*.json; do echo "$f: $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d.get('kernel_ms_per_loop'), (d.get('roofline') or {}).get('frac'))" 2>&1)"; done
