#!/bin/bash
# xhat mode (inner blocks store only the un-affined, pre-split LayerNorm output): learner parity suites, then A/B of DCC_TC_XHAT
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r02x}
timeout 900 python -m pytest tests/test_mappo_cuda.py tests/test_compact_cuda.py tests/test_rnn_cuda.py -m gpu -q --maxfail=12 -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/${TAG}_pytest.log | tail -14
for v in 1 0 1 0; do
DCC_TC_XHAT=$v timeout 300 python tools/bench_mappo.py --envs 8192 --iters 1 --epochs 4 --compact 1 > gpurun_out/${TAG}_mappo_xhat$v.log 2>&1
echo "xhat=$v: $(tail -2 gpurun_out/${TAG}_mappo_xhat$v.log | head -1 | cut -c1-100)"
done
