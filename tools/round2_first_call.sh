#!/bin/bash
# First GPU call of the next round: verify and measure what was built after round 1's GPU budget was spent
# (paired match-stream layout, camera alignment), in ONE gpurun invocation.  Everything lands in gpurun_out/.
#
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
#
# Reading the results:
#   exp_tests.log        gated GPU tests (tests/test_gpu_layout.py, tests/test_gpu_zz_align.py) -- must be green before anything else
#   bench_*_{plain,paired}.json   bench.py lines; compare kernel_ms_per_loop.ggs and roofline.frac between the layouts
#   probe_*_{plain,paired}.txt    per-stage cycles of the GGS iteration (stage1+2a is where the layout acts)
#   ggs_*_paired.ncu-rep          ncu --set full of the paired kernel (issue slots busy / DRAM throughput vs profiles/r1_final_summary.md)
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
export PDB_TEST_EXPERIMENTAL=1
timeout 900 python -m pytest tests/test_gpu_zz_align.py tests/test_gpu_layout.py -m gpu -q > gpurun_out/exp_tests.log 2>&1
echo "experimental tests exit code $?" | tee -a gpurun_out/exp_tests.log
for layout in plain paired; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --ggs-layout $layout > gpurun_out/bench_cfg3_$layout.json 2> gpurun_out/bench_cfg3_$layout.err
  timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload cfg5 --ggs-layout $layout > gpurun_out/bench_cfg5_$layout.json 2> gpurun_out/bench_cfg5_$layout.err
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --seqs-per-gpu 8 --ggs-layout $layout > gpurun_out/bench_b8_$layout.json 2> gpurun_out/bench_b8_$layout.err
  timeout 120 python tools/ggs_stage_probe.py 20 2048 $layout > gpurun_out/probe_cfg3_$layout.txt 2>&1
  timeout 300 python tools/ggs_stage_probe.py 80 4096 $layout > gpurun_out/probe_cfg5_$layout.txt 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ggs_entry -s 2 -c 1 -f -o gpurun_out/ggs_cfg5_paired \
  python tools/ggs_stage_probe.py 80 4096 paired > gpurun_out/ncu_cfg5_paired.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ggs_entry -s 2 -c 1 -f -o gpurun_out/ggs_cfg3_paired \
  python tools/ggs_stage_probe.py 20 2048 paired > gpurun_out/ncu_cfg3_paired.log 2>&1
# tensor-core engine at a size where the projections are real GEMMs (512 sequences x 20 frames = 10 240 tokens, GGS off):
# steps/s and the tensor-pipe share of tc_linear_kernel (sm__pipe_tensor_cycles_active) for DESIGN 4.3
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload cfg2 --seqs-per-gpu 512 > gpurun_out/bench_cfg2_b512.json 2> gpurun_out/bench_cfg2_b512.err
timeout 600 ncu --set full --clock-control none -k regex:tc_linear -s 200 -c 6 -f -o gpurun_out/tc_linear_b512 \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload cfg2 --seqs-per-gpu 512 > gpurun_out/ncu_tc_b512.log 2>&1
tail -n 3 gpurun_out/exp_tests.log
for f in gpurun_out/bench_*_*.json; do echo "$f: $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d.get('kernel_ms_per_loop'), (d.get('roofline') or {}).get('frac'))" 2>&1)"; done
