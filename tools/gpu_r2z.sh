#!/bin/bash
# generic A/B of one environment knob on the loop bench + learner parity suites.  usage: gpu_r2z.sh TAG KNOB [values...]
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r02z}; KNOB=${2:-DCC_LN_VEC}; shift 2 || true
VALS=${@:-1 0 1 0}
timeout 900 python -m pytest tests/test_mappo_cuda.py tests/test_compact_cuda.py tests/test_rnn_cuda.py -m gpu -q --maxfail=12 -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/${TAG}_pytest.log | tail -14
for v in $VALS; do
env $KNOB=$v timeout 300 python tools/bench_mappo.py --envs 8192 --iters 1 --epochs 4 --compact 1 > gpurun_out/${TAG}_mappo_$v.log 2>&1
echo "$KNOB=$v: $(tail -2 gpurun_out/${TAG}_mappo_$v.log | head -1 | cut -c1-100)"
done
