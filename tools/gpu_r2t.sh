#!/bin/bash
# critic super-chunks (N actor chunks per critic pass): learner parity suites, then A/B of DCC_CRITIC_SUPER at 8192 envs / 4 epochs
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r02t}
timeout 900 python -m pytest tests/test_mappo_cuda.py tests/test_compact_cuda.py -m gpu -q --maxfail=12 -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/${TAG}_pytest.log | tail -16
for v in 1 0 1 0; do
DCC_CRITIC_SUPER=$v timeout 300 python tools/bench_mappo.py --envs 8192 --iters 1 --epochs 4 --compact 1 > gpurun_out/${TAG}_mappo_super$v.log 2>&1
echo "super=$v: $(tail -2 gpurun_out/${TAG}_mappo_super$v.log | head -1 | cut -c1-130)"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${TAG}_launches_compact1.csv \
    python tools/bench_mappo.py --envs 65536 --T 5 --epochs 1 --iters 1 --compact 1 > gpurun_out/${TAG}_ncu_compact1.log 2>&1
python tools/agg_launches.py gpurun_out/${TAG}_launches_compact1.csv 40 > gpurun_out/${TAG}_launches_compact1.txt 2>&1
head -30 gpurun_out/${TAG}_launches_compact1.txt
