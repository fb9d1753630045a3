#!/bin/bash
# Round 2, GPU call D (1 GPU): --set full capture of one activation chunk of the update (DRAM traffic per chunk), configs[2]
# bench line + launch list + --set full capture of the 16/256 env kernel.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r02d}
# one chunk = 18 kernels; launch 460 is the first kernel (compact_features) of the first chunk of the 2nd iteration's update
timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 460 --launch-count 18 -f -o gpurun_out/${TAG}_chunk_full \
    python tools/bench_mappo.py --envs 65536 --T 4 --epochs 1 --iters 1 --compact 1 > gpurun_out/${TAG}_ncu_chunk.log 2>&1
ncu -i gpurun_out/${TAG}_chunk_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_chunk_raw.csv 2>/dev/null
python tools/chunk_traffic.py gpurun_out/${TAG}_chunk_raw.csv | tail -22
# configs[2]
( timeout 600 python bench.py --workload env16 ) > gpurun_out/${TAG}_bench_env16.json 2> gpurun_out/${TAG}_bench_env16.err; tail -c 1800 gpurun_out/${TAG}_bench_env16.json; tail -3 gpurun_out/${TAG}_bench_env16.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches_bench_env16.csv \
    python bench.py --workload env16 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench_env16.log 2>&1
python tools/agg_launches.py gpurun_out/${TAG}_launches_bench_env16.csv 8
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dcc_env_spec_kernel -s 8 -c 2 -f -o gpurun_out/${TAG}_env16_full \
    python bench.py --workload env16 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_env16_full.log 2>&1
ncu -i gpurun_out/${TAG}_env16_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_env16_raw.csv 2>/dev/null
python tools/chunk_traffic.py gpurun_out/${TAG}_env16_raw.csv | tail -4
ls -la gpurun_out/${TAG}_* | awk '{print $5, $9}'
