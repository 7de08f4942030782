#!/bin/bash
# A/B of the denoiser's stage hand-over (flag-carrying words vs group barriers) + the hand-over microbenchmark.
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
(timeout 120 build/stage_probe 20 64 20000; timeout 120 build/stage_probe 20 128 20000; timeout 120 build/stage_probe 8 64 20000) > gpurun_out/fl_stage_probe.txt 2>&1
cat gpurun_out/fl_stage_probe.txt
for f in 1 0; do
  PDB_DEN_FLAG=$f timeout 300 python tools/den_stage_probe.py 20 1 > gpurun_out/fl_probe_den_n20_flag$f.txt 2>&1
  tail -n 9 gpurun_out/fl_probe_den_n20_flag$f.txt
done
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/fl_gputests.log 2>&1
echo "gpu tests exit code $?" | tee -a gpurun_out/fl_gputests.log
tail -n 3 gpurun_out/fl_gputests.log
for f in 1 0; do
  PDB_DEN_FLAG=$f timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/fl_bench_cfg3_flag$f.json 2> gpurun_out/fl_bench_cfg3_flag$f.err
  PDB_DEN_FLAG=$f timeout 300 python bench.py --steps 10 --warmup 3 --workload cfg1 --no-cpu-baseline > gpurun_out/fl_bench_cfg1_flag$f.json 2> gpurun_out/fl_bench_cfg1_flag$f.err
  PDB_DEN_FLAG=$f timeout 300 python bench.py --steps 10 --warmup 3 --workload cfg2 --no-cpu-baseline > gpurun_out/fl_bench_cfg2_flag$f.json 2> gpurun_out/fl_bench_cfg2_flag$f.err
  PDB_DEN_FLAG=$f timeout 600 python bench.py --steps 3 --warmup 3 --workload cfg5 --no-cpu-baseline > gpurun_out/fl_bench_cfg5_flag$f.json 2> gpurun_out/fl_bench_cfg5_flag$f.err
  PDB_DEN_FLAG=$f timeout 600 python bench.py --steps 3 --warmup 3 --workload cfg4 --no-cpu-baseline --denoiser-engine fp32 > gpurun_out/fl_bench_cfg4_fp32_flag$f.json 2> gpurun_out/fl_bench_cfg4_fp32_flag$f.err
done
for f in gpurun_out/fl_bench_*.json; do echo "$f: $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['value'],1), d.get('kernel_ms_per_loop'), d.get('e2e',{}).get('value'))" 2>&1)"; done
