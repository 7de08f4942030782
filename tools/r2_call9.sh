#!/bin/bash
# Round 2, GPU call 9: single-warp stage 2a, idle warps skip the tail, evict-first stream policy, swap-AB opt-in.
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/gputests9.log 2>&1
echo "gpu tests exit code $?" | tee -a gpurun_out/gputests9.log
timeout 120 python tools/ggs_stage_probe.py 20 2048 > gpurun_out/probe9_cfg3.txt 2>&1
timeout 300 python tools/ggs_stage_probe.py 80 4096 > gpurun_out/probe9_cfg5.txt 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench9_cfg3.json 2> gpurun_out/bench9_cfg3.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload cfg4 > gpurun_out/bench9_cfg4_1gpu.json 2> gpurun_out/bench9_cfg4_1gpu.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload cfg5 > gpurun_out/bench9_cfg5.json 2> gpurun_out/bench9_cfg5.err
tail -n 4 gpurun_out/gputests9.log
cat gpurun_out/probe9_cfg3.txt gpurun_out/probe9_cfg5.txt
for f in gpurun_out/bench9_*.json; do echo "$f: $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d.get('kernel_ms_per_loop'), (d.get('roofline') or {}).get('frac'))" 2>&1)"; done
