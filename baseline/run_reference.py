#!/usr/bin/env python
"""Times the UNMODIFIED reference (zhaozijie2022/dynamic-coverage-control, Python/NumPy/torch) on this box's host cores.

    python baseline/run_reference.py env  --n 8 --m 64 --procs 16 --steps 150 --warmup 10
    python baseline/run_reference.py loop --iters 3            # shipped 4 UAV / 20 PoI config, 16 SubprocVecEnv workers

Measurement scaffolding for bench.py's `cpu_baseline.python_reference` — not
product code.  The reference sources are imported where they lie: /root/reference/uav_dcc_control in the build
container, else the verbatim copy under baseline/_ref/ (git-ignored, shipped to the GPU box by gpurun), behind the stub
shim for the plumbing modules it imports but this image lacks (tests/oracle_shim: gym / imp / omegaconf / imageio /
wandb).  Every arithmetic line that runs is the reference's own; the harness lifts only the 4 / 20 literals of
`Scenario.make_world` (tests/golden/ref_harness.py::GenScenario) so that 8 UAV / 64 PoI can run at all.
Prints ONE JSON object on the last line.
"""
import argparse
import json
import os
import sys
import time
from argparse import Namespace

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def find_reference():
    for cand in (os.environ.get("DCC_REFERENCE_ROOT"), "/root/reference/uav_dcc_control",
                 os.path.join(HERE, "_ref", "uav_dcc_control")):
        if cand and os.path.isdir(cand):
            return cand
    return None


def bench_env(a, ref_root):
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import ref_harness
    ref = ref_harness.load_reference()
    Gen = ref["GenScenario"]
    N, M, P = a.n, a.m, a.procs

    class SynthScenario(Gen):
        reference_compat = True            # shipped semantics: the scenario's comm arguments never reach the world

        def __init__(self, *args, **kw):
            super().__init__(*args, **kw)
            if a.layout == "synthetic":    # the layout bench.py's CUDA arm uses (uniform(-1,1), seed 0)
                self.pos_pois = np.random.default_rng(0).uniform(-1.0, 1.0, (self.num_pois, 2))

    class _Mod:
        Scenario = SynthScenario
    cwd = os.getcwd()
    os.chdir(ref_root)
    try:
        import envs.make_env as mk
        import envs.mpe.uav_dcc as uav
    finally:
        os.chdir(cwd)
    uav.scenarios = Namespace(load=lambda name: _Mod)
    cfg = Namespace(env_file="mpe.uav_dcc", env_class="DCEnv", scenario_name="coverage", num_agents=N, num_pois=M,
                    max_ep_len=150, r_cover=0.2, r_comm=0.4, comm_r_scale=0.95, comm_force_scale=0.0, seed=0,
                    n_rollout_threads=P)
    env = mk.make_env(cfg)                 # the reference's own factory: SubprocVecEnv, one OS process per env (P > 1)
    rng = np.random.default_rng(0)
    acts = [rng.standard_normal((P, N, 2)).astype(np.float32) for _ in range(4)]
    obs = env.reset()
    assert obs.shape == (P, N, 4 + 2 * (N - 1) + 5 * M), obs.shape
    for t in range(a.warmup):
        env.step([x.copy() for x in acts[t % 4]])
    t0 = time.perf_counter()
    for t in range(a.steps):
        obs, rew, done, infos = env.step([x.copy() for x in acts[t % 4]])
    dt = time.perf_counter() - t0
    env.close()
    return {"what": "env", "agent_steps_per_s": P * N * a.steps / dt, "env_steps_per_s": P * a.steps / dt, "seconds": dt,
            "n_agents": N, "n_pois": M, "procs": P, "steps": a.steps, "vec_env": type(env).__name__, "layout": a.layout}


def bench_loop(a, ref_root):
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import ref_harness
    ref_harness.load_reference()
    cwd = os.getcwd()
    os.chdir(ref_root)
    try:
        from omegaconf import OmegaConf
        import utils.pytorch_utils as ptu
        from learner import Learner
        cfg = OmegaConf.merge(OmegaConf.load("./config/env_config/dcc.yaml"),
                              OmegaConf.load("./config/algo_config/mappo.yaml"), OmegaConf.load("./config/expt.yaml"))
    finally:
        os.chdir(cwd)
    ptu.set_gpu_mode(False, 0)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    # the shipped configuration (4 UAV / 20 PoI, 16 SubprocVecEnv workers, T = 150, 15 PPO epochs); only logging,
    # checkpoints and the eval / render envs are switched off
    cfg.update(n_eval_rollout_threads=0, n_render_rollout_threads=0, log_wandb=False, save_model=False)
    if a.procs:
        cfg.update(n_rollout_threads=a.procs)
    lr = Learner(cfg)
    E, N, T = int(cfg.n_rollout_threads), int(cfg.num_agents), int(cfg.max_ep_len)
    t_roll = t_upd = 0.0
    lr.warmup(lr.rl_buffer, lr.train_envs)
    for it in range(1, a.iters + 1):
        lr.trainer.policy.lr_decay(it, cfg.n_iters)
        t0 = time.perf_counter()
        lr.rollout(lr.rl_buffer, lr.train_envs)
        t1 = time.perf_counter()
        lr.rl_update()
        t2 = time.perf_counter()
        t_roll += t1 - t0
        t_upd += t2 - t1
    lr.train_envs.close()
    return {"what": "loop", "agent_steps_per_s": E * N * T * a.iters / (t_roll + t_upd), "rollout_s_per_iter": t_roll / a.iters,
            "update_s_per_iter": t_upd / a.iters, "n_agents": N, "n_pois": int(cfg.num_pois), "procs": E, "T": T,
            "ppo_epoch": int(cfg.ppo_epoch), "iters": a.iters, "torch_threads": threads}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["env", "loop"])
    ap.add_argument("--n", type=int, default=8)
    ap.add_argument("--m", type=int, default=64)
    ap.add_argument("--procs", type=int, default=0)
    ap.add_argument("--steps", type=int, default=150)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--layout", default="synthetic", choices=["synthetic", "reference"])
    a = ap.parse_args()
    ref_root = find_reference()
    if ref_root is None:
        print(json.dumps({"unavailable": "reference sources not found (/root/reference, baseline/_ref)"}))
        return
    os.environ["DCC_REFERENCE_ROOT"] = ref_root
    if a.what == "env" and not a.procs:
        a.procs = os.cpu_count() or 1
    out = bench_env(a, ref_root) if a.what == "env" else bench_loop(a, ref_root)
    import numpy
    import torch
    out.update(cpu_count=os.cpu_count(), reference_root=ref_root, numpy=numpy.__version__, torch=torch.__version__)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
