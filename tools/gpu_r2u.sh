#!/bin/bash
# recurrent policies: the new GPU suite, then the whole GPU suite (ACT_IDENT / float64 ln0_finalize touched shared kernels)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r02u}
timeout 900 python -m pytest tests/test_rnn_cuda.py -m gpu -q --maxfail=6 -p no:cacheprovider > gpurun_out/${TAG}_pytest_rnn.log 2>&1
echo "rnn pytest exit $?"; tail -25 gpurun_out/${TAG}_pytest_rnn.log | cut -c1-220
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider --deselect tests/test_rnn_cuda.py > gpurun_out/${TAG}_pytest_all.log 2>&1
echo "all pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${TAG}_pytest_all.log | tail -12
timeout 300 python tools/bench_mappo.py --envs 8192 --iters 1 --epochs 4 --compact 1 > gpurun_out/${TAG}_mappo.log 2>&1
tail -2 gpurun_out/${TAG}_mappo.log | cut -c1-200
