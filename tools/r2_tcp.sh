cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
(python tools/tc_gemm_probe.py 160; PDB_TC_NO_SMALL=1 python tools/tc_gemm_probe.py 160; python tools/tc_gemm_probe.py 640) 2>&1 | tee gpurun_out/tc_gemm_probe.txt
