#!/bin/bash
# CTAs per sequence of the GGS kernel at config 3 (PDB_GGS_CPP caps the plan of csrc/api_core.cu:ggs_plan).
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
for c in 148 111 74 37; do
  echo "==== PDB_GGS_CPP=$c"
  PDB_GGS_CPP=$c timeout 120 python tools/ggs_stage_probe.py 20 2048 2>&1 | grep -E "launch|cycles/iter"
  PDB_GGS_CPP=$c timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> /dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', round(d['value'],1), d.get('kernel_ms_per_loop'), (d.get('roofline') or {}).get('frac'))"
done 2>&1 | tee gpurun_out/sweep_cpp_cfg3.txt
