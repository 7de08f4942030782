#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_features.py -m gpu -q > gpurun_out/tc_gputests.log 2>&1
echo "gpu tests exit code $?" | tee -a gpurun_out/tc_gputests.log
tail -n 3 gpurun_out/tc_gputests.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/tc_launches_b8.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workload cfg2 --seqs-per-gpu 8 > gpurun_out/tc_launches_b8.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/tc_launches_b8.csv') if not l.startswith('==')))
hdr = rows[0]
ki, vi, gi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size')
for r in rows[1:][40:100]:
    print(r[ki][:50], r[gi], r[vi])
PY
