#!/bin/bash
# Round 2, GPU call 3: exchange microbenchmark v3 (one-hop vector-reduction all-reduce, polling variants), the three GPU tests
# whose tolerances / operating point were fixed, and the denoiser stage probe.
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
timeout 300 build/xchg_probe 148 20000 > gpurun_out/xchg_probe3.txt 2>&1
timeout 300 build/xchg_probe 565 5000 > gpurun_out/xchg_probe3_n80.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_layout.py -m gpu -q -k "full_loop or seeded_sizes" > gpurun_out/gputests3.log 2>&1
echo "gpu tests exit code $?" | tee -a gpurun_out/gputests3.log
timeout 300 python tools/den_stage_probe.py 20 1 > gpurun_out/den_probe_n20.txt 2>&1
timeout 300 python tools/den_stage_probe.py 80 1 > gpurun_out/den_probe_n80.txt 2>&1
cat gpurun_out/xchg_probe3.txt; cat gpurun_out/xchg_probe3_n80.txt | tail -n +7
tail -n 5 gpurun_out/gputests3.log
cat gpurun_out/den_probe_n20.txt gpurun_out/den_probe_n80.txt
