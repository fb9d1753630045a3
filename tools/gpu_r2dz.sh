#!/bin/bash
# pre-split gradients (DCC_TC_DZSPLIT): whole GPU suite, A/B at 8192 envs, launch list of the update
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r02dz2}
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/${TAG}_pytest_all.log 2>&1
echo "all pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${TAG}_pytest_all.log | tail -12
for v in 1 0 1 0; do
DCC_TC_DZSPLIT=$v timeout 300 python tools/bench_mappo.py --envs 8192 --iters 1 --epochs 4 --compact 1 > gpurun_out/${TAG}_mappo_$v.log 2>&1
echo "DCC_TC_DZSPLIT=$v: $(tail -2 gpurun_out/${TAG}_mappo_$v.log | head -1 | cut -c1-200)"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python tools/bench_mappo.py --envs 65536 --T 5 --epochs 1 --iters 1 --compact 1 > gpurun_out/${TAG}_ncu.log 2>&1
python tools/agg_launches.py gpurun_out/${TAG}_launches.csv 24 > gpurun_out/${TAG}_launches.txt 2>&1
head -16 gpurun_out/${TAG}_launches.txt
