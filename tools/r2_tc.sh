#!/bin/bash
# Tensor-core engine at a few hundred tokens: deep ring + split-K + in-kernel LayerNorm statistics vs the previous launcher.
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_features.py -m gpu -q -x > gpurun_out/tc_gputests.log 2>&1
echo "gpu tests exit code $?" | tee -a gpurun_out/tc_gputests.log
tail -n 3 gpurun_out/tc_gputests.log
for v in new old; do
  if [ $v = old ]; then export PDB_TC_NO_SMALL=1; else unset PDB_TC_NO_SMALL; fi
  timeout 600 python bench.py --steps 3 --warmup 3 --workload cfg4 --no-cpu-baseline > gpurun_out/tc_bench_cfg4_$v.json 2> gpurun_out/tc_bench_cfg4_$v.err
  timeout 600 python bench.py --steps 3 --warmup 3 --workload cfg2 --seqs-per-gpu 32 --no-cpu-baseline > gpurun_out/tc_bench_b32_$v.json 2> gpurun_out/tc_bench_b32_$v.err
  timeout 300 python bench.py --steps 3 --warmup 3 --workload features --no-cpu-baseline > gpurun_out/tc_bench_features_$v.json 2> gpurun_out/tc_bench_features_$v.err
done
unset PDB_TC_NO_SMALL
for f in gpurun_out/tc_bench_*.json; do echo "$f: $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['value'],1), d.get('kernel_ms_per_loop'), d.get('e2e',{}).get('value'))" 2>&1)"; done
