#!/bin/bash
# weight-gradient kernel in isolation: fp16 split with fp32 operands / X pre-split (TMA) / X and dZ pre-split (TMA, no producer warps)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-tcw}
{
for nout in 256 160 288; do
  for split in 0 1 2; do
    echo -n "R=303104 Nout=$nout split=$split: "
    timeout 30 tools/tc_bench_np.bin wgrad 303104 $nout 1 $split | tail -1
  done
done
} > gpurun_out/${TAG}_wgrad.log 2>&1
cat gpurun_out/${TAG}_wgrad.log | cut -c1-200
