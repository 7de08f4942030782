#!/bin/bash
# compute-sanitizer over the round-2 kernels (GPU box only): memcheck on the GGS / Sampson / sampler entry points, both denoiser
# hand-overs and the tensor-core launcher regimes (split-K, in-kernel LayerNorm statistics, programmatic dependent launch);
# racecheck (shared-memory hazards) on the persistent GGS and denoiser kernels.  Small cases only: the tools slow a launch 10-100x.
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
timeout 1500 $S --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
  -k "sampson_eval_vs_reference or five_phases or early_exit or batch_equals_singles or host_buffer_entry or host_matches_entry or (handover and (1-5 or 3-13)) or wrong_problem_count" \
  > gpurun_out/san_memcheck.log 2>&1
echo "memcheck exit $?" | tee -a gpurun_out/san_memcheck.log
timeout 1200 $S --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_tc.py -m gpu -q -x \
  -k "in_place or (engine and tiles128 and (8-20 or 7-20 or 1-5 or 32-20)) or (matches_fp32 and tiles128 and plain)" \
  > gpurun_out/san_tc_memcheck.log 2>&1
echo "tc memcheck exit $?" | tee -a gpurun_out/san_tc_memcheck.log
timeout 1500 $S --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
  -k "early_exit or (five_phases) or (handover and 1-5) or (denoiser_forward_vs_reference)" \
  > gpurun_out/san_racecheck.log 2>&1
echo "racecheck exit $?" | tee -a gpurun_out/san_racecheck.log
for f in gpurun_out/san_*.log; do echo "== $f"; grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|exit" $f | tail -n 4; done
