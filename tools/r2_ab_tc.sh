#!/bin/bash
# A/B of the tensor-core engine launcher regimes: new (split-K + in-kernel statistics, with / without programmatic dependent launch) vs the previous launcher (PDB_TC_NO_SMALL=1).
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_features.py -m gpu -q > gpurun_out/tc_gputests.log 2>&1
echo "gpu tests exit code $?" | tee -a gpurun_out/tc_gputests.log
tail -n 3 gpurun_out/tc_gputests.log
for v in pdl nopdl old; do
  unset PDB_TC_NO_SMALL PDB_TC_PDL
  if [ $v = old ]; then export PDB_TC_NO_SMALL=1 PDB_TC_PDL=0; fi
  if [ $v = nopdl ]; then export PDB_TC_PDL=0; fi
  timeout 600 python bench.py --steps 3 --warmup 3 --workload cfg4 --no-cpu-baseline > gpurun_out/tc_bench_cfg4_$v.json 2> gpurun_out/tc_bench_cfg4_$v.err
  timeout 600 python bench.py --steps 3 --warmup 3 --workload cfg2 --seqs-per-gpu 32 --no-cpu-baseline > gpurun_out/tc_bench_b32_$v.json 2> gpurun_out/tc_bench_b32_$v.err
  timeout 600 python bench.py --steps 3 --warmup 3 --workload cfg2 --seqs-per-gpu 128 --no-cpu-baseline > gpurun_out/tc_bench_b128_$v.json 2> gpurun_out/tc_bench_b128_$v.err
done
unset PDB_TC_NO_SMALL PDB_TC_PDL
timeout 300 python bench.py --steps 3 --warmup 3 --workload features --no-cpu-baseline > gpurun_out/tc_bench_features.json 2> gpurun_out/tc_bench_features.err
for f in gpurun_out/tc_bench_*.json; do echo "$f: $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['value'],1), d.get('kernel_ms_per_loop'), d.get('e2e',{}).get('value'))" 2>&1)"; done
