#!/bin/bash
# ncu passes for profiles/: launch lists (time only) and --set full captures of the dominant kernels.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r01f}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench_env.csv \
    python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_ncu_bench_env.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dcc_env_spec_kernel -s 8 -c 2 -f -o gpurun_out/${TAG}_env_full \
    python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_ncu_env_full.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG}_launches_mappo.csv \
    python tools/bench_mappo.py --envs 2048 --iters 0 --epochs 2 > gpurun_out/${TAG}_ncu_mappo.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 330 -c 8 -f -o gpurun_out/${TAG}_tc_full \
    python tools/bench_mappo.py --envs 2048 --iters 0 --epochs 1 > gpurun_out/${TAG}_ncu_tc_full.log 2>&1
ls -la gpurun_out/ | tail -12
