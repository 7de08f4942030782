#!/bin/bash
# BASELINE configs[3]: 64 sequences x 20 frames, GGS on, sharded over 8 B200 (8 sequences per GPU), both denoiser engines.
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 900 -- 'bash tools/r2_cfg4_8gpu.sh'
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
run() {  # $1 = tag, rest = bench arguments
  tag=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 8 --workload cfg4 --steps 3 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/cfg4_8gpu_$tag.json 2> gpurun_out/cfg4_8gpu_$tag.err
  echo "$tag exit $?"; tail -n 1 gpurun_out/cfg4_8gpu_$tag.json | cut -c1-600
}
run tf32
run fp32 --denoiser-engine fp32
