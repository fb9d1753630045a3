#!/bin/bash
# compute-sanitizer over the kernels added after r02s: recurrent path (dcc_rnn.cuh), xhat mode (relu_lnx_bwd_pipe_kernel, folded inner
# blocks), critic super-chunks, the vectorised LayerNorm-backward column map, the rewritten compact feature kernel
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r02s2}
SEL='test_recurrent_learner_vs_reference_golden and (h256 or rn2) and not 40 or recurrent_learner_end_to_end or compact_update_equals or compact_learner_vs_reference_golden and (gen_8x64_h256 or mb2) or test_learner_vs_reference_golden and (gen_8x64_h256 or net_layer2 or net_layer3) and not rnn'
timeout 1700 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_rnn_cuda.py tests/test_compact_cuda.py tests/test_mappo_cuda.py -m gpu -q -x -p no:cacheprovider -k "$SEL" > gpurun_out/${TAG}_memcheck.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" gpurun_out/${TAG}_memcheck.log | tail -8
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_compact_cuda.py tests/test_rnn_cuda.py -m gpu -q -x -p no:cacheprovider -k "compact_update_equals and 4-20 or test_recurrent_learner_vs_reference_golden and chunk_4x20_h32 and not 40" > gpurun_out/${TAG}_racecheck.log 2>&1
echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard|Error" gpurun_out/${TAG}_racecheck.log | tail -8
