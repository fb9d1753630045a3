#!/bin/bash
# ncu --set full over ONE super-chunk of the update (N actor chunks + one critic pass) with xhat mode; launch list of the update
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r02y}
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off --launch-count 120 -f -o gpurun_out/${TAG}_superchunk_full \
    python tools/bench_mappo.py --envs 65536 --T 5 --epochs 1 --iters 1 --compact 1 --profile-update > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/${TAG}_ncu.log | cut -c1-200
ncu -i gpurun_out/${TAG}_superchunk_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_superchunk_raw.csv 2>/dev/null
python tools/chunk_traffic.py gpurun_out/${TAG}_superchunk_raw.csv gpurun_out/${TAG}_superchunk_traffic.json first=compact_features count=88 > gpurun_out/${TAG}_superchunk_kernels.txt
tail -30 gpurun_out/${TAG}_superchunk_kernels.txt
ls -la gpurun_out/${TAG}_superchunk_full.ncu-rep
