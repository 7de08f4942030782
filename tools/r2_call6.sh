#!/bin/bash
# Round 2, GPU call 6: ncu source-level capture of the GGS kernel at config 3 (paired layout), probes, denoiser check.
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_layout.py -m gpu -q -x -k "ggs or sampson or loop" > gpurun_out/gputests6.log 2>&1
echo "gpu tests exit code $?" | tee -a gpurun_out/gputests6.log
timeout 120 python tools/ggs_stage_probe.py 20 2048 paired > gpurun_out/probe6_cfg3_paired.txt 2>&1
timeout 300 python tools/den_stage_probe.py 20 1 > gpurun_out/den_probe6_n20.txt 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --ggs-layout paired > gpurun_out/bench6_cfg3_paired.json 2> gpurun_out/bench6_cfg3_paired.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ggs_entry -s 2 -c 1 -f -o gpurun_out/ggs6_cfg3_paired \
  python tools/ggs_stage_probe.py 20 2048 paired > gpurun_out/ncu6_cfg3_paired.log 2>&1
ncu -i gpurun_out/ggs6_cfg3_paired.ncu-rep --page raw --csv > gpurun_out/ggs6_cfg3_paired_raw.csv 2>/dev/null
ncu -i gpurun_out/ggs6_cfg3_paired.ncu-rep --page source --csv > gpurun_out/ggs6_cfg3_paired_source.csv 2>/dev/null
tail -n 4 gpurun_out/gputests6.log
cat gpurun_out/probe6_cfg3_paired.txt gpurun_out/den_probe6_n20.txt
for f in gpurun_out/bench6_*.json; do echo "$f: $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d.get('kernel_ms_per_loop'), (d.get('roofline') or {}).get('frac'))" 2>&1)"; done
