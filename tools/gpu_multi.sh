#!/bin/bash
# Multi-GPU checks (gpurun --gpus N): sharded-update parity, env bench and MAPPO-loop bench under torchrun.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-2}; TAG=${2:-r01j}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29533 tools/dist_check.py > gpurun_out/${TAG}_dist_check_${N}gpu.log 2>&1; echo "dist_check exit $?"
grep '"ok"' gpurun_out/${TAG}_dist_check_${N}gpu.log | cut -c1-400
timeout 600 $TR --master-port 29534 bench.py --gpus $N > gpurun_out/${TAG}_bench_env_${N}gpu.json 2> gpurun_out/${TAG}_bench_env_${N}gpu.err; tail -c 700 gpurun_out/${TAG}_bench_env_${N}gpu.json
timeout 900 $TR --master-port 29535 bench.py --gpus $N --workload mappo --envs 4096 --steps 2 --warmup 3 > gpurun_out/${TAG}_bench_mappo_${N}gpu.json 2> gpurun_out/${TAG}_bench_mappo_${N}gpu.err; tail -c 900 gpurun_out/${TAG}_bench_mappo_${N}gpu.json
timeout 300 $TR --master-port 29536 bench.py --gpus $N --impl reference --steps 3 --warmup 1 | tail -c 400
