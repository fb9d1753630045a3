"""Learning-curve sanity run of the re-hosted trainer (shipped 4 UAV / 20 PoI hyper-parameters, many envs):
prints the reference's log lines; reward and coverage_rate must go up.
Usage: python tools/train_sanity.py [iters] [envs] [key=value ...]   (extra config overrides, e.g. num_agents=8 num_pois=64
per_env_layouts=True reference_compat=False comm_force_scale=1.0 num_mini_batch=2)"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from dcc_b200.learner import Learner  # noqa: E402
from dcc_b200.utils.config import load_config  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 60
envs = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
over = {}
for kv in sys.argv[3:]:
    k, v = kv.split("=", 1)
    over[k] = {"true": True, "false": False}.get(v.lower(), None)
    if over[k] is None:
        try:
            over[k] = int(v)
        except ValueError:
            try:
                over[k] = float(v)
            except ValueError:
                over[k] = v
cfg = load_config(None, n_rollout_threads=envs, n_iters=iters, n_eval_rollout_threads=0, save_model=False,
                  pos_pois_path=None, **over)
lr = Learner(cfg)
hist = []
t0 = time.time()
for it in range(1, iters + 1):
    lr.policy.lr_decay(it, cfg.n_iters)
    ri = lr.rollout(lr.rl_buffer, lr.train_envs)
    ti = lr.rl_update()
    hist.append((ri["reward"], ri["coverage_rate"], ti["value_loss"], ti["dist_entropy"], ri["connect_rate"]))
    if it % 5 == 0 or it == 1:
        lr.log(iter_=it, rollout_info=ri, rl_train_info=ti)
h = np.array(hist)
out = {"iters": iters, "envs": envs, "seconds": time.time() - t0, "reward_first5": float(h[:5, 0].mean()),
       "reward_last5": float(h[-5:, 0].mean()), "coverage_first5": float(h[:5, 1].mean()), "coverage_last5": float(h[-5:, 1].mean()),
       "connect_first5": float(h[:5, 4].mean()), "connect_last5": float(h[-5:, 4].mean()),
       "agent_steps": lr.agent_steps, "backend": lr.policy.gemm_backend(), "overrides": over}
print(json.dumps(out))
assert out["reward_last5"] > out["reward_first5"] and out["coverage_last5"] >= out["coverage_first5"], "no learning signal"
