#!/bin/bash
# compute-sanitizer over the learner / compact GPU tests with the round-2 kernels (TMA-fed operands, row pipeline, fused heads)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r02s}
SEL='compact_update_equals or obs_from_state or compact_learner_vs_reference_golden and (gen_8x64_h256 or mb2 or ship) or gemm_primitive or update_vs_oracle_random_batch or test_learner_vs_reference_golden and (gen_8x64_h256 or net_layer2)'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_compact_cuda.py tests/test_mappo_cuda.py -m gpu -q -x -p no:cacheprovider -k "$SEL" > gpurun_out/${TAG}_memcheck.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" gpurun_out/${TAG}_memcheck.log | tail -8
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_compact_cuda.py -m gpu -q -x -p no:cacheprovider -k "compact_update_equals and 4-20" > gpurun_out/${TAG}_racecheck.log 2>&1
echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard|Error" gpurun_out/${TAG}_racecheck.log | tail -8
