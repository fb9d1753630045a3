#!/bin/bash
# 2-GPU checks of the final build: sharded update == single-GPU big batch (tools/dist_check.py), a recurrent policy trained under
# torchrun (gradient all-reduce per minibatch, equal collective counts on both ranks), default bench line on 2 GPUs
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r02fin}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tools/dist_check.py > gpurun_out/${TAG}_dist_check_2gpu.log 2>&1
tail -3 gpurun_out/${TAG}_dist_check_2gpu.log | cut -c1-700
timeout 600 $TR --master-port 29513 -m dcc_b200.train 0 none num_agents=4 num_pois=20 n_rollout_threads=256 n_iters=6 use_recurrent_policy=True \
    n_eval_rollout_threads=0 n_render_rollout_threads=0 save_model=False log_wandb=False > gpurun_out/${TAG}_train_rnn_2gpu.log 2>&1
echo "rnn 2-gpu train exit $?"; grep -E "iter: |rollout_info" gpurun_out/${TAG}_train_rnn_2gpu.log | tail -4 | cut -c1-200
timeout 900 $TR --master-port 29512 bench.py --gpus 2 > gpurun_out/${TAG}_bench_default_2gpu.json 2> gpurun_out/${TAG}_bench_default_2gpu.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_default_2gpu.json").read().strip().splitlines()[-1])
f=d["full_loop"]; print("2 GPUs: env", d["value"], "full_loop", f["value"], "update_ms", f["update_ms"], "allreduce_ms", f["allreduce_ms"])
PY
