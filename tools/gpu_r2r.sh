#!/bin/bash
# compact minibatch path + learning sanity with the final kernels
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r02r}
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/${TAG}_pytest.log | tail -12
timeout 600 python tools/train_sanity.py 60 1024 > gpurun_out/${TAG}_train_sanity_4x20_1024envs.log 2>&1; tail -1 gpurun_out/${TAG}_train_sanity_4x20_1024envs.log | cut -c1-400
timeout 600 python tools/train_sanity.py 24 1024 num_agents=8 num_pois=64 reference_compat=False comm_force_scale=1.0 num_mini_batch=2 > gpurun_out/${TAG}_train_sanity_8x64_force_mb2.log 2>&1; tail -1 gpurun_out/${TAG}_train_sanity_8x64_force_mb2.log | cut -c1-500
