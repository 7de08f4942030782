#!/bin/bash
# Round 2 evidence run (one GPU): every GPU test, the bench lines of all single-GPU workloads, stage probes, and the ncu
# captures (--set full, raw CSV export) of the dominant kernels.  Outputs land in gpurun_out/ev_*; copy into profiles/.
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/ev_gputests.log 2>&1
echo "gpu tests exit code $?" | tee -a gpurun_out/ev_gputests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/ev_bench_cfg3.json 2> gpurun_out/ev_bench_cfg3.err
timeout 300 python bench.py --steps 10 --warmup 3 --workload cfg1 --no-cpu-baseline > gpurun_out/ev_bench_cfg1.json 2> gpurun_out/ev_bench_cfg1.err
timeout 300 python bench.py --steps 10 --warmup 3 --workload cfg2 --no-cpu-baseline > gpurun_out/ev_bench_cfg2.json 2> gpurun_out/ev_bench_cfg2.err
timeout 600 python bench.py --steps 3 --warmup 3 --workload cfg5 --no-cpu-baseline > gpurun_out/ev_bench_cfg5.json 2> gpurun_out/ev_bench_cfg5.err
timeout 600 python bench.py --steps 3 --warmup 3 --workload cfg4 --no-cpu-baseline > gpurun_out/ev_bench_cfg4_1gpu_tf32.json 2> gpurun_out/ev_bench_cfg4_1gpu_tf32.err
timeout 600 python bench.py --steps 3 --warmup 3 --workload cfg4 --no-cpu-baseline --denoiser-engine fp32 > gpurun_out/ev_bench_cfg4_1gpu_fp32.json 2> gpurun_out/ev_bench_cfg4_1gpu_fp32.err
timeout 300 python bench.py --steps 5 --warmup 3 --workload features --no-cpu-baseline > gpurun_out/ev_bench_features.json 2> gpurun_out/ev_bench_features.err
timeout 120 python tools/ggs_stage_probe.py 20 2048 > gpurun_out/ev_probe_ggs_cfg3.txt 2>&1
timeout 300 python tools/ggs_stage_probe.py 80 4096 > gpurun_out/ev_probe_ggs_cfg5.txt 2>&1
timeout 300 python tools/den_stage_probe.py 20 1 > gpurun_out/ev_probe_den_n20.txt 2>&1
timeout 200 python tools/tc_gemm_probe.py 160 > gpurun_out/ev_tc_gemm_probe.txt 2>&1
if [ -x build/stage_probe ]; then
  (timeout 120 build/stage_probe 20 64 20000; timeout 120 build/stage_probe 20 128 20000; timeout 120 build/stage_probe 8 64 20000) > gpurun_out/ev_stage_probe.txt 2>&1
fi
# launch list of one headline loop (duration-only pass)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ev_launches_cfg3.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ev_launches_cfg3.log 2>&1
# --set full captures
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ggs_entry -s 2 -c 1 -f -o gpurun_out/ev_ggs_cfg3 python tools/ggs_stage_probe.py 20 2048 > gpurun_out/ev_ncu_ggs_cfg3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ggs_entry -s 2 -c 1 -f -o gpurun_out/ev_ggs_cfg5 python tools/ggs_stage_probe.py 80 4096 > gpurun_out/ev_ncu_ggs_cfg5.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:denoiser_kernel -s 1 -c 1 -f -o gpurun_out/ev_den_n20 python tools/den_stage_probe.py 20 1 > gpurun_out/ev_ncu_den_n20.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:tc_linear -s 100 -c 4 -f -o gpurun_out/ev_tc_b8 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workload cfg2 --seqs-per-gpu 8 > gpurun_out/ev_ncu_tc_b8.log 2>&1
for n in ev_ggs_cfg3 ev_ggs_cfg5 ev_den_n20 ev_tc_b8; do
  ncu -i gpurun_out/$n.ncu-rep --page raw --csv > gpurun_out/${n}_raw.csv 2>/dev/null
done
rm -f gpurun_out/ev_den_n20.ncu-rep gpurun_out/ev_tc_b8.ncu-rep
tail -n 3 gpurun_out/ev_gputests.log
for f in gpurun_out/ev_bench_*.json; do echo "$f: $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['value'],1), d.get('kernel_ms_per_loop'), (d.get('roofline') or {}).get('frac'), d.get('e2e',{}).get('value'))" 2>&1)"; done
cat gpurun_out/ev_probe_ggs_cfg3.txt
