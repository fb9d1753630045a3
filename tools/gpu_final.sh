#!/bin/bash
# Final validation of a round inside one short gpurun call: all GPU parity tests, smoke, both own-arm bench lines,
# a learning-curve sanity run and one ncu capture of the fp16-split forward GEMM.  Outputs -> gpurun_out/.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-final}
timeout 170 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/${TAG}_smoke.log
timeout 200 python bench.py --workload mappo > gpurun_out/${TAG}_bench_mappo.json 2> gpurun_out/${TAG}_bench_mappo.err; tail -c 1500 gpurun_out/${TAG}_bench_mappo.json
timeout 120 python bench.py > gpurun_out/${TAG}_bench_env.json 2> gpurun_out/${TAG}_bench_env.err; tail -c 1200 gpurun_out/${TAG}_bench_env.json
timeout 100 python tools/train_sanity.py 60 1024 > gpurun_out/${TAG}_train_sanity_4x20_1024envs.log 2>&1; tail -1 gpurun_out/${TAG}_train_sanity_4x20_1024envs.log | cut -c1-400
timeout 70 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_fwd --launch-skip 3 --launch-count 1 -f -o gpurun_out/${TAG}_ncu_fwd_f16 tools/tc_bench_np.bin 303104 352 1 0 1 1 2 > gpurun_out/${TAG}_ncu.log 2>&1; tail -2 gpurun_out/${TAG}_ncu.log
