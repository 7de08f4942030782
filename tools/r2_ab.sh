#!/bin/bash
# A/B run: GGS tests + benches of the three GGS workloads + stage probe (used for single changes of the GGS kernel).
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_layout.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/ab_gputests.log 2>&1
echo "gpu tests exit code $?" | tee -a gpurun_out/ab_gputests.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ab_bench_cfg3.json 2> gpurun_out/ab_bench_cfg3.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload cfg5 > gpurun_out/ab_bench_cfg5.json 2> gpurun_out/ab_bench_cfg5.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload cfg4 > gpurun_out/ab_bench_cfg4_1gpu.json 2> gpurun_out/ab_bench_cfg4_1gpu.err
timeout 120 python tools/ggs_stage_probe.py 20 2048 > gpurun_out/ab_probe_cfg3.txt 2>&1
timeout 300 python tools/ggs_stage_probe.py 80 4096 > gpurun_out/ab_probe_cfg5.txt 2>&1
tail -n 3 gpurun_out/ab_gputests.log
for f in gpurun_out/ab_bench_*.json; do echo "$f: $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d.get('kernel_ms_per_loop'), (d.get('roofline') or {}).get('frac'))" 2>&1)"; done
cat gpurun_out/ab_probe_cfg3.txt gpurun_out/ab_probe_cfg5.txt
