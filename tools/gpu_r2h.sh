#!/bin/bash
# epilogue rework check: learner parity suites, role profile, loop timing
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r02h}
timeout 900 python -m pytest tests/test_mappo_cuda.py tests/test_compact_cuda.py -m gpu -q --maxfail=15 -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed|exit|Error" gpurun_out/${TAG}_pytest.log | tail -20
bash tools/gpu_r2g.sh ${TAG} | grep -v "^==" | cut -c1-330
for i in 1 2; do
timeout 300 python tools/bench_mappo.py --envs 8192 --iters 1 --epochs 4 --compact 1 > gpurun_out/${TAG}_mappo_$i.log 2>&1
echo "loop: $(tail -2 gpurun_out/${TAG}_mappo_$i.log | head -1 | cut -c1-130)"
done
