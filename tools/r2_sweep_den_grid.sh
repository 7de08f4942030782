#!/bin/bash
# CTAs of the persistent fp32 denoiser kernel at 20 tokens (PDB_DEN_GRID caps the grid of csrc/api_sampler.cu:enqueue_denoiser).
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
for g in 148 128 96 64; do
  echo "==== PDB_DEN_GRID=$g"
  PDB_DEN_GRID=$g timeout 200 python bench.py --steps 10 --warmup 3 --workload cfg2 --no-cpu-baseline 2> /dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', round(d['value'],1), d.get('kernel_ms_per_loop'))"
done 2>&1 | tee gpurun_out/sweep_den_grid_cfg2.txt
