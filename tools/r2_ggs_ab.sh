#!/bin/bash
# GGS change check: parity tests that touch the GGS kernel, stage probe, headline benches.
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_layout.py -m gpu -q -x > gpurun_out/gg_gputests.log 2>&1
echo "gpu tests exit code $?" | tee -a gpurun_out/gg_gputests.log
tail -n 3 gpurun_out/gg_gputests.log
timeout 120 python tools/ggs_stage_probe.py 20 2048 > gpurun_out/gg_probe_cfg3.txt 2>&1
timeout 300 python tools/ggs_stage_probe.py 80 4096 > gpurun_out/gg_probe_cfg5.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/gg_bench_cfg3.json 2> gpurun_out/gg_bench_cfg3.err
timeout 600 python bench.py --steps 3 --warmup 3 --workload cfg5 --no-cpu-baseline > gpurun_out/gg_bench_cfg5.json 2> gpurun_out/gg_bench_cfg5.err
timeout 600 python bench.py --steps 3 --warmup 3 --workload cfg4 --no-cpu-baseline > gpurun_out/gg_bench_cfg4.json 2> gpurun_out/gg_bench_cfg4.err
for f in gpurun_out/gg_bench_*.json; do echo "$f: $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['value'],1), d.get('kernel_ms_per_loop'), (d.get('roofline') or {}).get('frac'), d.get('e2e',{}).get('value'))" 2>&1)"; done
cat gpurun_out/gg_probe_cfg3.txt; tail -n 9 gpurun_out/gg_probe_cfg5.txt
