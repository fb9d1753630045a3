#!/bin/bash
# 8-GPU records (gpurun --gpus 8): env bench and full MAPPO loop (BASELINE configs[4]) under torchrun.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-8}; TAG=${2:-r01m}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29534 bench.py --gpus $N --no-cpu-baseline > gpurun_out/${TAG}_bench_env_${N}gpu.json 2> gpurun_out/${TAG}_bench_env_${N}gpu.err; tail -c 600 gpurun_out/${TAG}_bench_env_${N}gpu.json
timeout 900 $TR --master-port 29535 bench.py --gpus $N --workload mappo --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_mappo_${N}gpu.json 2> gpurun_out/${TAG}_bench_mappo_${N}gpu.err; tail -c 900 gpurun_out/${TAG}_bench_mappo_${N}gpu.json
