#!/bin/bash
# head LayerNorm-backward ring depth A/B (DCC_HEAD_SLOTS=3 vs 6), parity of the learner suites with the default
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r02w}
timeout 900 python -m pytest tests/test_mappo_cuda.py tests/test_compact_cuda.py -m gpu -q --maxfail=12 -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/${TAG}_pytest.log | tail -8
for v in 6 3 6 3; do
DCC_HEAD_SLOTS=$v timeout 300 python tools/bench_mappo.py --envs 8192 --iters 1 --epochs 4 --compact 1 > gpurun_out/${TAG}_mappo_slots$v.log 2>&1
echo "slots=$v: $(tail -2 gpurun_out/${TAG}_mappo_slots$v.log | head -1 | cut -c1-100)"
done
