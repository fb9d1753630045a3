#!/bin/bash
# final validation of round 2 (second session): whole GPU suite, smoke, learning sanity (MLP 4/20, 8/64 with force + minibatches,
# recurrent 4/20), configs[2] bench line, reference arm
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r02fin}
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/${TAG}_pytest.log | tail -12
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/${TAG}_smoke.log
timeout 600 python tools/train_sanity.py 60 1024 > gpurun_out/${TAG}_train_sanity_4x20_1024envs.log 2>&1; tail -1 gpurun_out/${TAG}_train_sanity_4x20_1024envs.log | cut -c1-400
timeout 600 python tools/train_sanity.py 24 1024 num_agents=8 num_pois=64 reference_compat=False comm_force_scale=1.0 num_mini_batch=2 > gpurun_out/${TAG}_train_sanity_8x64_force_mb2.log 2>&1; tail -1 gpurun_out/${TAG}_train_sanity_8x64_force_mb2.log | cut -c1-500
timeout 900 python tools/train_sanity.py 40 256 use_recurrent_policy=True > gpurun_out/${TAG}_train_sanity_4x20_rnn_256envs.log 2>&1; tail -1 gpurun_out/${TAG}_train_sanity_4x20_rnn_256envs.log | cut -c1-400
timeout 300 python bench.py --workload env16 > gpurun_out/${TAG}_bench_env16_1gpu.json 2> gpurun_out/${TAG}_bench_env16.err; tail -c 600 gpurun_out/${TAG}_bench_env16_1gpu.json
timeout 600 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_bench_reference.err; tail -c 900 gpurun_out/${TAG}_bench_reference_arm.json
