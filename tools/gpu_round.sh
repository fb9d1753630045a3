#!/bin/bash
# One gpurun call: GPU parity tests, smoke, chunk A/B of the MAPPO update, both bench arms.  Outputs -> gpurun_out/.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r01d}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider --durations=15 > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/${TAG}_smoke.log
for CH in 37888; do
  timeout 300 python tools/bench_mappo.py --envs 8192 --iters 1 --epochs 4 --chunk $CH > gpurun_out/${TAG}_mappo_chunk${CH}.log 2>&1
  tail -2 gpurun_out/${TAG}_mappo_chunk${CH}.log
done
timeout 600 python bench.py --impl reference > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; tail -c 600 gpurun_out/${TAG}_bench_ref.json
timeout 600 python bench.py > gpurun_out/${TAG}_bench_env.json 2> gpurun_out/${TAG}_bench_env.err; tail -c 1500 gpurun_out/${TAG}_bench_env.json
timeout 900 python bench.py --workload mappo > gpurun_out/${TAG}_bench_mappo.json 2> gpurun_out/${TAG}_bench_mappo.err; tail -c 1500 gpurun_out/${TAG}_bench_mappo.json
timeout 600 python tools/train_sanity.py 60 1024 > gpurun_out/${TAG}_train_sanity_4x20_1024envs.log 2>&1; tail -1 gpurun_out/${TAG}_train_sanity_4x20_1024envs.log | cut -c1-400
timeout 600 python tools/train_sanity.py 24 1024 num_agents=8 num_pois=64 reference_compat=False comm_force_scale=1.0 num_mini_batch=2 > gpurun_out/${TAG}_train_sanity_8x64_force_mb2.log 2>&1; tail -1 gpurun_out/${TAG}_train_sanity_8x64_force_mb2.log | cut -c1-500
