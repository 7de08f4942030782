#!/bin/bash
# Round 2, GPU call 10: production GGS kernels without probe code, static touched set in stage 2b.
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_layout.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/gputests10.log 2>&1
echo "gpu tests exit code $?" | tee -a gpurun_out/gputests10.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench10_cfg3.json 2> gpurun_out/bench10_cfg3.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload cfg5 > gpurun_out/bench10_cfg5.json 2> gpurun_out/bench10_cfg5.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload cfg4 > gpurun_out/bench10_cfg4_1gpu.json 2> gpurun_out/bench10_cfg4_1gpu.err
timeout 120 python tools/ggs_stage_probe.py 20 2048 > gpurun_out/probe10_cfg3.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,sm__icc_request_hit_rate.pct,sm__icc_requests.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:ggs_entry -s 3 -c 2 --csv --log-file gpurun_out/icc10_cfg3.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/icc10_cfg3.log 2>&1
tail -n 3 gpurun_out/gputests10.log
for f in gpurun_out/bench10_*.json; do echo "$f: $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d.get('kernel_ms_per_loop'), (d.get('roofline') or {}).get('frac'))" 2>&1)"; done
cat gpurun_out/probe10_cfg3.txt; tail -n 12 gpurun_out/icc10_cfg3.csv
