#!/bin/bash
# First GPU check of the experimental fp16-split weight-gradient kernel (DCC_TC_WGRAD_F16=1): isolated timing against
# the 3xTF32 kernel, the GEMM primitive test with the dW shapes, the learner parity tests and an in-box loop A/B.
# Every step runs under its own timeout: the kernel has never run on hardware.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-wg16}
{
for shape in "303104 338" "303104 256" "37888 2704"; do
  for f in 0 1; do timeout 30 tools/tc_bench.bin wgrad $shape $f | head -3; done
done
} > gpurun_out/${TAG}_tcbench.log 2>&1
grep "^wgrad" gpurun_out/${TAG}_tcbench.log
DCC_TC_WGRAD_F16=1 timeout 300 python -m pytest tests/test_mappo_cuda.py -m gpu -q --maxfail=10 -p no:cacheprovider -k "gemm_primitive" > gpurun_out/${TAG}_pytest_gemm.log 2>&1
tail -3 gpurun_out/${TAG}_pytest_gemm.log
DCC_TC_WGRAD_F16=1 timeout 600 python -m pytest tests/test_mappo_cuda.py -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed|exit" gpurun_out/${TAG}_pytest.log | tail -15
for f in 0 1 0 1; do
  DCC_TC_WGRAD_F16=$f timeout 300 python tools/bench_mappo.py --envs 8192 --iters 1 --epochs 4 > gpurun_out/${TAG}_mappo_wg$f.log 2>&1
  echo "DCC_TC_WGRAD_F16=$f: $(tail -2 gpurun_out/${TAG}_mappo_wg$f.log | head -1 | cut -c1-120)"
done
