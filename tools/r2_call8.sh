#!/bin/bash
# Round 2, GPU call 8: tensor-core engine below 128 tokens (swap-AB vs 128-token tiles vs fp32 kernel), ncu of the swap-AB
# launches, finer stage-3 probe, compute-sanitizer on the round-2 kernels.
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q > gpurun_out/gputests8_tc.log 2>&1
echo "tc tests exit code $?" | tee -a gpurun_out/gputests8_tc.log
timeout 300 python tools/tc_small_probe.py > gpurun_out/tc_small_swap.json 2> gpurun_out/tc_small_swap.err
PDB_TC_SWAP=0 timeout 300 python tools/tc_small_probe.py > gpurun_out/tc_small_noswap.json 2> gpurun_out/tc_small_noswap.err
timeout 600 ncu --set full --clock-control none -k regex:tc_linear -s 70 -c 6 -f -o gpurun_out/tc_swap_n20 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workload cfg2 --denoiser-engine tf32 > gpurun_out/ncu8_tc_swap.log 2>&1
ncu -i gpurun_out/tc_swap_n20.ncu-rep --page raw --csv > gpurun_out/tc_swap_n20_raw.csv 2>/dev/null; rm -f gpurun_out/tc_swap_n20.ncu-rep
timeout 120 python tools/ggs_stage_probe.py 20 2048 > gpurun_out/probe8_cfg3.txt 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ggs_five or sampson_eval_vs_reference or ggs_batch or fused_loop_rejects" > gpurun_out/sanitizer8_memcheck.log 2>&1
echo "memcheck exit $?" | tee -a gpurun_out/sanitizer8_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ggs_five or ggs_batch" > gpurun_out/sanitizer8_racecheck.log 2>&1
echo "racecheck exit $?" | tee -a gpurun_out/sanitizer8_racecheck.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k "tc_linear_matches" > gpurun_out/sanitizer8_tc_memcheck.log 2>&1
echo "tc memcheck exit $?" | tee -a gpurun_out/sanitizer8_tc_memcheck.log
tail -n 3 gpurun_out/gputests8_tc.log; cat gpurun_out/tc_small_swap.json gpurun_out/tc_small_noswap.json; cat gpurun_out/probe8_cfg3.txt
tail -n 4 gpurun_out/sanitizer8_memcheck.log gpurun_out/sanitizer8_racecheck.log gpurun_out/sanitizer8_tc_memcheck.log
