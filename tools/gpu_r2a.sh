#!/bin/bash
# Round 2, GPU call A: parity suite (incl. the compact-state path and the 8/64/H=256 golden), first hardware run of the
# fp16-split weight-gradient / dX kernels, compact vs materialised loop A/B, per-kernel launch lists.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r02a}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider --durations=10 > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed|exit" gpurun_out/${TAG}_pytest.log | tail -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/${TAG}_smoke.log
for C in 1 0; do
  timeout 300 python tools/bench_mappo.py --envs 8192 --iters 1 --epochs 4 --compact $C > gpurun_out/${TAG}_mappo_compact$C.log 2>&1
  echo "compact=$C: $(tail -2 gpurun_out/${TAG}_mappo_compact$C.log | head -1 | cut -c1-140)"
done
# experimental fp16-split backward kernels: isolated timing, primitive test, learner parity, loop A/B
{
for shape in "303104 338" "303104 256" "303104 160" "37888 2704" "37888 288"; do
  for f in 0 1; do timeout 30 tools/tc_bench.bin wgrad $shape $f | head -3; done
done
} > gpurun_out/${TAG}_tcbench.log 2>&1
grep "^wgrad" gpurun_out/${TAG}_tcbench.log
DCC_TC_WGRAD_F16=1 timeout 300 python -m pytest tests/test_mappo_cuda.py -m gpu -q --maxfail=10 -p no:cacheprovider -k "gemm_primitive" > gpurun_out/${TAG}_wg16_pytest_gemm.log 2>&1
tail -3 gpurun_out/${TAG}_wg16_pytest_gemm.log
DCC_TC_WGRAD_F16=1 timeout 900 python -m pytest tests/test_mappo_cuda.py tests/test_compact_cuda.py -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/${TAG}_wg16_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_wg16_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed|exit" gpurun_out/${TAG}_wg16_pytest.log | tail -25
DCC_TC_WGRAD_F16=1 timeout 300 python tools/bench_mappo.py --envs 8192 --iters 1 --epochs 4 --compact 1 > gpurun_out/${TAG}_mappo_compact1_wg16.log 2>&1
echo "compact=1 wgrad_f16=1: $(tail -2 gpurun_out/${TAG}_mappo_compact1_wg16.log | head -1 | cut -c1-140)"
# per-kernel launch lists of one update pass over 6.9 chunks (65536 envs x T = 4)
for C in 1 0; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches_compact$C.csv \
      python tools/bench_mappo.py --envs 65536 --T 4 --epochs 1 --iters 1 --compact $C > gpurun_out/${TAG}_ncu_compact$C.log 2>&1
  python tools/agg_launches.py gpurun_out/${TAG}_launches_compact$C.csv 40 > gpurun_out/${TAG}_launches_compact$C.txt 2>&1
  head -16 gpurun_out/${TAG}_launches_compact$C.txt
done
