"""Full MAPPO loop timing (BASELINE configs[3]: rollout + GAE + ppo_epoch-epoch update) on one GPU; tuning aid.
    python tools/bench_mappo.py --envs 4096 --T 150 --iters 2 [--backend 1|2] [--epochs 15]
Prints per-phase device times (CUDA events) and agent-steps/s of the whole iteration."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from dcc_b200.learner import Learner  # noqa: E402
from dcc_b200.utils.config import load_config  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8)
    ap.add_argument("--m", type=int, default=64)
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--T", type=int, default=150)
    ap.add_argument("--epochs", type=int, default=15)
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--backend", type=int, default=0)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--compact", type=int, default=-1, help="1 / 0 = compact / materialised rollout storage (default: auto)")
    ap.add_argument("--profile-update", action="store_true",
                    help="cudaProfilerStart/Stop around the LAST iteration's update (ncu --profile-from-start off)")
    a = ap.parse_args()
    cfg = load_config(None, num_agents=a.n, num_pois=a.m, n_rollout_threads=a.envs, max_ep_len=a.T, ppo_epoch=a.epochs,
                      n_iters=a.iters + 1, n_eval_rollout_threads=0, save_model=False, gemm_backend=a.backend,
                      chunk_rows=a.chunk, poi_layout="synthetic", compact_rollout=None if a.compact < 0 else bool(a.compact))
    lr = Learner(cfg)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    out = []
    for it in range(1, a.iters + 2):
        lr.policy.lr_decay(it, cfg.n_iters)
        ev[0].record()
        ri = lr.rollout(lr.rl_buffer, lr.train_envs)
        ev[1].record()
        prof = a.profile_update and it == a.iters + 1
        if prof:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
        ti = lr.rl_update()
        if prof:
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
        ev[2].record()
        torch.cuda.synchronize()
        r_ms, u_ms = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
        steps = a.envs * a.n * a.T
        out.append(dict(iter=it, rollout_ms=r_ms, update_ms=u_ms, agent_steps_per_s=steps / ((r_ms + u_ms) * 1e-3),
                        rollout_info=ri, value_loss=ti["value_loss"], ratio=ti["ratio"]))
        print(json.dumps(out[-1]), flush=True)
    print(json.dumps(dict(backend=lr.policy.gemm_backend(), compact=lr.compact, chunk_rows=lr.policy.lib.dcc_mappo_chunk_rows(lr.policy._h),
                          launches=lr.policy.launch_count())))


if __name__ == "__main__":
    main()
