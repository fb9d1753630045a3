#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r02i}
{
for bin in tools/tc_bench_sl1.bin tools/tc_bench_sl0.bin tools/tc_bench_np.bin; do
  echo "== $bin"
  timeout 30 $bin 303104 160 1 0 1 1 2 0 1 1
  timeout 30 $bin 303104 256 1 0 1 1 2 2 0 1
  timeout 30 $bin 303104 256 1 0 1 1 2 2 0 0
  timeout 30 $bin 303104 256 0 0 1 1 2 0 1 1
done
} > gpurun_out/${TAG}_tcbench_fwd.log 2>&1
cat gpurun_out/${TAG}_tcbench_fwd.log | cut -c1-330
timeout 600 python -m pytest tests/test_mappo_cuda.py tests/test_compact_cuda.py -m gpu -q --maxfail=15 -p no:cacheprovider -x > gpurun_out/${TAG}_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/${TAG}_pytest.log | tail -8
timeout 300 python tools/bench_mappo.py --envs 8192 --iters 1 --epochs 4 --compact 1 > gpurun_out/${TAG}_mappo_1.log 2>&1
echo "loop: $(tail -2 gpurun_out/${TAG}_mappo_1.log | head -1 | cut -c1-130)"
