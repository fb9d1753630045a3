#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r02k}
{
for bin in tools/tc_bench.bin tools/tc_bench_np.bin; do
  echo "== $bin"
  timeout 30 $bin 303104 160 1 0 1 1 2 0 1 1
  timeout 30 $bin 303104 256 1 0 1 1 2 2 0 1
  timeout 30 $bin 303104 256 1 0 1 1 2 2 0 0
  timeout 30 $bin 303104 256 0 0 1 1 2 0 1 1
  timeout 30 $bin 37888 288 1 0 1 1 2 0 1 1
done
} > gpurun_out/${TAG}_tcbench.log 2>&1
cat gpurun_out/${TAG}_tcbench.log | cut -c1-330
timeout 900 python -m pytest tests/test_mappo_cuda.py tests/test_compact_cuda.py -m gpu -q --maxfail=15 -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/${TAG}_pytest.log | tail -8
for i in 1 2; do
timeout 300 python tools/bench_mappo.py --envs 8192 --iters 1 --epochs 4 --compact 1 > gpurun_out/${TAG}_mappo_$i.log 2>&1
echo "loop: $(tail -2 gpurun_out/${TAG}_mappo_$i.log | head -1 | cut -c1-130)"
done
