#!/bin/bash
# role profile of the forward-shaped tcgen05 kernel after the round-2 changes (finer epilogue timers)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r02g}
{
for bin in tools/tc_bench.bin tools/tc_bench_np.bin; do
  echo "== $bin"
  #                          M      K  epi dbg tma f16 pf head h c
  timeout 30 $bin 303104 160 1 0 1 1 2 0 1 1
  timeout 30 $bin 303104 256 1 0 1 1 2 2 0 1
  timeout 30 $bin 303104 256 1 0 1 1 2 2 0 0
  timeout 30 $bin 303104 256 1 0 1 1 2 0 1 1
  timeout 30 $bin 303104 256 0 0 1 1 2 0 1 1
  timeout 30 $bin 303104 256 1 1 1 1 2 2 0 1
  timeout 30 $bin 37888 288 1 0 1 1 2 0 1 1
done
} > gpurun_out/${TAG}_tcbench_fwd.log 2>&1
cat gpurun_out/${TAG}_tcbench_fwd.log
