#!/bin/bash
# ncu --set full + source counters of ONE launch of a kernel inside the loop bench.  usage: gpu_ncu_one.sh TAG KERNEL_REGEX [skip]
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-one}; K=${2:-compact_features}; SKIP=${3:-6}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip $SKIP --launch-count 1 -f -o gpurun_out/${TAG} \
    python tools/bench_mappo.py --envs 65536 --T 3 --epochs 1 --iters 1 --compact 1 > gpurun_out/${TAG}.log 2>&1
echo "ncu exit $?"
ls -la gpurun_out/${TAG}.ncu-rep
ncu -i gpurun_out/${TAG}.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py 2>/dev/null | head -40
