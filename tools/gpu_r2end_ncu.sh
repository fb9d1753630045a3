#!/bin/bash
# ncu --set full over ONE actor chunk of the update on the final build (9 kernels), without source import (small report)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r02end}
timeout 900 ncu --set full --clock-control none --profile-from-start off --launch-count 40 -f -o /tmp/${TAG}_chunk \
    python tools/bench_mappo.py --envs 65536 --T 5 --epochs 1 --iters 1 --compact 1 --profile-update > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu exit $?"; ls -la /tmp/${TAG}_chunk.ncu-rep
ncu -i /tmp/${TAG}_chunk.ncu-rep --page raw --csv > /tmp/${TAG}_chunk_raw.csv 2>/dev/null
python tools/chunk_traffic.py /tmp/${TAG}_chunk_raw.csv gpurun_out/${TAG}_chunk_traffic.json first=compact_features count=9 > gpurun_out/${TAG}_ncu_chunk_kernels.txt
cat gpurun_out/${TAG}_ncu_chunk_kernels.txt
ncu -i /tmp/${TAG}_chunk.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > gpurun_out/${TAG}_ncu_chunk_full.txt 2>/dev/null
wc -c gpurun_out/${TAG}_ncu_chunk_full.txt
