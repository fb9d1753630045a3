#!/bin/bash
# Short gpurun call: GPU parity tests + the MAPPO-loop tuning bench.  usage: gpu_quick.sh TAG [pytest -k expr]
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-q}
KEXPR=${2:-}
if [ -n "$KEXPR" ]; then
  timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider -k "$KEXPR" > gpurun_out/${TAG}_pytest.log 2>&1
else
  timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
fi
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed|exit" gpurun_out/${TAG}_pytest.log | tail -15
timeout 300 python tools/bench_mappo.py --envs 8192 --iters 1 --epochs 4 > gpurun_out/${TAG}_mappo.log 2>&1
tail -2 gpurun_out/${TAG}_mappo.log
