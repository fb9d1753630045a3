#!/bin/bash
# refresh of the chunk traffic capture with the pre-split operand path + parity suites + default bench line
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r02p}; SKIP=${2:-482}
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/${TAG}_pytest.log | tail -12
timeout 900 ncu --set full --clock-control none --import-source on --launch-skip $SKIP --launch-count 18 -f -o gpurun_out/${TAG}_chunk_full \
    python tools/bench_mappo.py --envs 65536 --T 4 --epochs 1 --iters 1 --compact 1 > gpurun_out/${TAG}_ncu_chunk.log 2>&1
ncu -i gpurun_out/${TAG}_chunk_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_chunk_raw.csv 2>/dev/null
python tools/chunk_traffic.py gpurun_out/${TAG}_chunk_raw.csv | tail -22
( time timeout 900 python bench.py ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 1200 gpurun_out/${TAG}_bench.json; tail -4 gpurun_out/${TAG}_bench.err
