#!/bin/bash
# role profile + timing of the forward-shaped tcgen05 kernel in its product configurations; old = HEAD's kernel, new = working tree
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-tcb}
{
for bin in tools/tc_bench_old.bin tools/tc_bench.bin tools/tc_bench_old_np.bin tools/tc_bench_np.bin; do
  [ -x $bin ] || continue
  echo "== $bin"
  #                M      K  epi dbg tma f16 pf head sh sc as hs unit
  timeout 30 $bin 303104 160 1   0   1   1   0  0    1  0  1  1  1     # L1: xhat mode (no a store, unit affine, pre-split in/out)
  timeout 30 $bin 303104 256 1   0   1   1   0  2    0  1  1  0  0     # L2: fused head, stores a, pre-split in
  timeout 30 $bin 303104 256 0   0   1   1   2  0    1  1  0  0  0     # dX-shaped: raw store, fp32 A through the producer warps
done
} > gpurun_out/${TAG}_tcbench.log 2>&1
cat gpurun_out/${TAG}_tcbench.log | cut -c1-330
