#!/bin/bash
# Round 2, GPU call B: parity suite with the fused heads / pipelined LayerNorm backward / fp16-split dX, knob A/Bs of the
# loop, per-kernel launch list, and the new default bench line (full_loop + Python reference).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r02b}
timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider --durations=8 > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed|exit" gpurun_out/${TAG}_pytest.log | tail -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/${TAG}_smoke.log
run() {  # label, env assignments...
  local label=$1; shift
  env "$@" timeout 300 python tools/bench_mappo.py --envs 8192 --iters 1 --epochs 4 --compact 1 > gpurun_out/${TAG}_mappo_${label}.log 2>&1
  echo "${label}: $(tail -2 gpurun_out/${TAG}_mappo_${label}.log | head -1 | cut -c1-130)"
}
run all_on DCC_X=1
run head_off DCC_TC_HEAD=0
run pipe_off DCC_LN_PIPE=0
run dx16_off DCC_TC_DX_F16=0
run all_off DCC_TC_HEAD=0 DCC_LN_PIPE=0 DCC_TC_DX_F16=0
run all_on2 DCC_X=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches_compact1.csv \
    python tools/bench_mappo.py --envs 65536 --T 4 --epochs 1 --iters 1 --compact 1 > gpurun_out/${TAG}_ncu_compact1.log 2>&1
python tools/agg_launches.py gpurun_out/${TAG}_launches_compact1.csv 40 > gpurun_out/${TAG}_launches_compact1.txt 2>&1
head -14 gpurun_out/${TAG}_launches_compact1.txt
( time timeout 900 python bench.py ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 3000 gpurun_out/${TAG}_bench.json; tail -4 gpurun_out/${TAG}_bench.err
( time timeout 600 python bench.py --impl reference ) > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; tail -c 1500 gpurun_out/${TAG}_bench_ref.json; tail -4 gpurun_out/${TAG}_bench_ref.err
