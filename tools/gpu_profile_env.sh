#!/bin/bash
# ncu passes for the env bench: launch list (time only) + --set full capture of the step kernel.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r01p}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench_env.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench_env.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dcc_env_spec_kernel -s 8 -c 2 -f -o gpurun_out/${TAG}_env_full \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_env_full.log 2>&1
ls -la gpurun_out | tail -5
