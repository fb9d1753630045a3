#!/bin/bash
# Round 2, multi-GPU call: sharded update == single-GPU big batch (tools/dist_check.py), then the default bench line under torchrun
# (env step + full MAPPO loop with the NCCL gradient all-reduce, configs[4]).  Usage: gpu_r2e_multi.sh TAG NGPUS
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r02e}; N=${2:-2}
if [ "$N" = "2" ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/${TAG}_dist_check_${N}gpu.log 2>&1
  tail -3 gpurun_out/${TAG}_dist_check_${N}gpu.log | cut -c1-600
fi
( time NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N ) \
    > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
tail -c 2500 gpurun_out/${TAG}_bench_${N}gpu.json; tail -5 gpurun_out/${TAG}_bench_${N}gpu.err
