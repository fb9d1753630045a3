#!/bin/bash
# A/B of fp16-split forward GEMM variants inside ONE box (tc_bench_old.bin = the previous commit's kernel).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-f16}
{
for shape in "303104 352 1" "303104 256 1" "303104 256 0" "37888 2720 1"; do
  for bin in tc_bench_old tc_bench; do
    [ -x tools/$bin.bin ] || continue
    for cfg in "1 0" "1 2" "1 3"; do echo -n "$bin "; timeout 60 tools/$bin.bin $shape 0 1 $cfg | head -3; done
  done
done
} > gpurun_out/${TAG}_tcbench.log 2>&1
grep "fwd" gpurun_out/${TAG}_tcbench.log
timeout 600 python -m pytest tests/test_mappo_cuda.py -m gpu -q --maxfail=10 -p no:cacheprovider ${2:+-k "$2"} > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed|exit" gpurun_out/${TAG}_pytest.log | tail -15
for v in "0 0" "1 2"; do
  set -- $v
  DCC_TC_F16=$1 DCC_TC_PF=$2 timeout 300 python tools/bench_mappo.py --envs 8192 --iters 1 --epochs 4 > gpurun_out/${TAG}_mappo_f$1_pf$2.log 2>&1
  echo "DCC_TC_F16=$1 DCC_TC_PF=$2: $(tail -2 gpurun_out/${TAG}_mappo_f$1_pf$2.log | head -1 | cut -c1-120)"
done
