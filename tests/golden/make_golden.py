"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference env
(/root/reference/uav_dcc_control, via ref_harness.py) on seeded inputs.

    python tests/golden/make_golden.py            # writes tests/golden/env_*.npz

The reference ships no tests, fixtures or known-answer vectors, so these files are the
pin for the oracle (oracle/dcc_env_oracle.c) and for the CUDA path.  The env has no randomness
(core.py:79 `u_noise=None`; reset is fixed, scenarios/coverage.py:64-78), so the files are
reproducible bit for bit given numpy's BLAS ddot rounding (see oracle header).

Per case the file holds, for T steps of ONE env instance driven by recorded float32 actions:
  actions (T,N,2) f32 | pos_vel (T,N,4) f64 post-step, post-auto-reset | pos_vel_pre (T,N,4) f64 pre-reset
  energy (T,M) u8 post-auto-reset | energy_pre (T,M) u8 | reward (T,) f64 | done (T,) bool
  coverage_rate (T,) f64 | connect, connect_ (T,) bool | adj, adj_ (T,N) u32 row bitmasks
  obs_steps (K,) int32, obs (K,N,D) f32 = float32(reference obs after the wrapper's auto-reset)
  obs0 (N,D) f32 = reset observation | poi (M,2) f64 | cfg = json string
"unit" cases instead hold K independent (state, action) -> one-step results from injected states.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_harness import RefEnv, load_reference  # noqa: E402


def pack_rows(mat):
    n = mat.shape[0]
    out = np.zeros(n, dtype=np.uint32)
    for a in range(n):
        for b in range(n):
            if mat[a, b]:
                out[a] |= np.uint32(1 << b)
    return out


def policy_random(rng, sigma=1.0):
    def f(env, t):
        return rng.standard_normal((env.n_agents, 2)).astype(np.float32) * np.float32(sigma)
    return f


def policy_seek(rng, noise=0.3):
    """Greedy coverage controller: UAV i heads for the (i mod #undone)-th nearest undone PoI."""
    def f(env, t):
        w = env.world
        act = np.zeros((env.n_agents, 2), dtype=np.float32)
        undone = [lm for lm in w.landmarks if not lm.done]
        for i, ag in enumerate(w.agents):
            if undone:
                d = [np.linalg.norm(lm.state.p_pos - ag.state.p_pos) for lm in undone]
                order = np.argsort(d)
                tgt = undone[order[i % len(undone)]].state.p_pos
                v = tgt - ag.state.p_pos
                nv = np.linalg.norm(v)
                if nv > 1e-9:
                    v = v / nv
                act[i] = (v * 1.0).astype(np.float32)
            act[i] += (rng.standard_normal(2) * noise).astype(np.float32)
        return act
    return f


def policy_runaway(rng):
    """Push every UAV outward along a fixed heading: soft-bound penalty then |p|>1.5 termination."""
    def f(env, t):
        n = env.n_agents
        ang = np.arange(n) * (2 * np.pi / n) + 0.1
        act = np.stack([np.cos(ang), np.sin(ang)], 1) * (0.6 + 0.4 * (np.arange(n) % 3))[:, None]
        return (act + rng.standard_normal((n, 2)) * 0.05).astype(np.float32)
    return f


def run_traj(name, N, M, T, policy, seed, comm_r_scale=0.95, comm_force_scale=0.0, compat=False,
             obs_every=1, poi=None, r_cover=0.2, r_comm=0.4):
    rng = np.random.RandomState(seed)
    env = RefEnv(N, M, r_cover=r_cover, r_comm=r_comm, comm_r_scale=comm_r_scale, comm_force_scale=comm_force_scale,
                 reference_compat=compat, pos_pois=poi)
    pol = {"random": policy_random, "seek": policy_seek, "runaway": policy_runaway}[policy](rng)
    obs0 = env.reset()
    D = obs0.shape[1]
    rec = {k: [] for k in ("actions", "pos_vel", "pos_vel_pre", "energy", "energy_pre", "reward", "done",
                           "coverage_rate", "connect", "connect_", "adj", "adj_")}
    obs_steps, obs = [], []
    for t in range(T):
        a = pol(env, t)
        r = env.step(a)
        assert np.all(r["reward"] == r["reward"][0]) and np.all(r["done"] == r["done"][0])
        rec["actions"].append(a)
        rec["pos_vel"].append(r["pos_vel"]); rec["pos_vel_pre"].append(r["pos_vel_pre"])
        rec["energy"].append(r["energy"].astype(np.uint8)); rec["energy_pre"].append(r["energy_pre"].astype(np.uint8))
        rec["reward"].append(r["reward"][0]); rec["done"].append(r["done"][0])
        rec["coverage_rate"].append(r["coverage_rate"])
        rec["connect"].append(r["connect"]); rec["connect_"].append(r["connect_"])
        rec["adj"].append(pack_rows(r["adj"])); rec["adj_"].append(pack_rows(r["adj_"]))
        if t < 3 or t % obs_every == 0 or r["done"][0] or t == T - 1:
            obs_steps.append(t); obs.append(r["obs"].astype(np.float32))
    world = env.world
    cfg = dict(name=name, kind="traj", n_agents=N, n_pois=M, T=T, policy=policy, seed=seed, obs_dim=D,
               r_cover=r_cover, r_comm=r_comm,
               comm_r_scale=float(world.comm_r_scale), contact_force=float(world.contact_force),
               reference_compat=bool(compat))
    out = {k: np.array(v) for k, v in rec.items()}
    out.update(obs_steps=np.array(obs_steps, dtype=np.int32), obs=np.array(obs), obs0=obs0.astype(np.float32),
               poi=np.array(env.scenario.pos_pois, dtype=np.float64), cfg=np.array(json.dumps(cfg)))
    path = os.path.join(HERE, "env_%s.npz" % name)
    np.savez_compressed(path, **out)
    nd = int(out["done"].sum())
    print("%-28s T=%d D=%d dones=%d connect_false=%d connect__false=%d maxcov=%.2f  %.0f KB" % (
        name, T, D, nd, int((~out["connect"]).sum()), int((~out["connect_"]).sum()),
        out["coverage_rate"].max(), os.path.getsize(path) / 1024))


def run_unit(name, N, M, K, seed, comm_r_scale=0.95, comm_force_scale=0.0, compat=False, spread=1.0, r_cover=0.2,
             r_comm=0.4):
    """K independent one-step transitions from injected random states."""
    rng = np.random.RandomState(seed)
    env = RefEnv(N, M, r_cover=r_cover, r_comm=r_comm, comm_r_scale=comm_r_scale, comm_force_scale=comm_force_scale,
                 reference_compat=compat)
    env.reset()
    keys = ("actions", "pos_vel_in", "energy_in", "pos_vel", "pos_vel_pre", "energy", "energy_pre", "reward", "done",
            "coverage_rate", "connect", "connect_", "adj", "adj_", "obs")
    rec = {k: [] for k in keys}
    for k in range(K):
        mode = k % 4
        pv = np.zeros((N, 4))
        if mode == 0:    # clustered near a random centre (mostly connected)
            c = rng.uniform(-0.8, 0.8, 2)
            pv[:, 0:2] = c + rng.uniform(-0.45, 0.45, (N, 2)) * spread
        elif mode == 1:  # spread over the arena (mostly disconnected)
            pv[:, 0:2] = rng.uniform(-1.2, 1.2, (N, 2))
        elif mode == 2:  # near / beyond the bounds
            pv[:, 0:2] = rng.uniform(-1.6, 1.6, (N, 2))
        else:            # chain with spacing near the comm thresholds (0.76 .. 0.8)
            step = rng.uniform(0.70, 0.84, N)
            ang = rng.uniform(0, 2 * np.pi)
            pv[:, 0] = -1.0 + np.cumsum(step) * np.cos(ang) * 0.5
            pv[:, 1] = np.cumsum(step) * np.sin(ang) * 0.5
            pv[:, 0:2] = np.clip(pv[:, 0:2], -1.45, 1.45)
        pv[:, 2:4] = rng.uniform(-0.4, 0.4, (N, 2))
        en = rng.randint(0, 9, M).astype(np.float64)
        if k % 7 == 0:
            en[:] = np.where(rng.rand(M) < 0.9, 6.0, 4.0)  # nearly all done: exercise the all-done branch
        if k % 11 == 0:
            en[:] = 7.0                                      # already all done
        env.set_state(pv, en)
        a = (rng.standard_normal((N, 2)) * (1.0 if k % 3 else 3.0)).astype(np.float32)
        r = env.step(a)
        rec["actions"].append(a); rec["pos_vel_in"].append(pv); rec["energy_in"].append(en.astype(np.uint8))
        rec["pos_vel"].append(r["pos_vel"]); rec["pos_vel_pre"].append(r["pos_vel_pre"])
        rec["energy"].append(r["energy"].astype(np.uint8)); rec["energy_pre"].append(r["energy_pre"].astype(np.uint8))
        rec["reward"].append(r["reward"][0]); rec["done"].append(r["done"][0])
        rec["coverage_rate"].append(r["coverage_rate"])
        rec["connect"].append(r["connect"]); rec["connect_"].append(r["connect_"])
        rec["adj"].append(pack_rows(r["adj"])); rec["adj_"].append(pack_rows(r["adj_"]))
        rec["obs"].append(r["obs"].astype(np.float32))
    world = env.world
    D = rec["obs"][0].shape[1]
    cfg = dict(name=name, kind="unit", n_agents=N, n_pois=M, K=K, seed=seed, obs_dim=D, r_cover=r_cover, r_comm=r_comm,
               comm_r_scale=float(world.comm_r_scale), contact_force=float(world.contact_force),
               reference_compat=bool(compat))
    out = {k: np.array(v) for k, v in rec.items()}
    out.update(poi=np.array(env.scenario.pos_pois, dtype=np.float64), cfg=np.array(json.dumps(cfg)))
    path = os.path.join(HERE, "env_%s.npz" % name)
    np.savez_compressed(path, **out)
    print("%-28s K=%d D=%d dones=%d connect_false=%d connect__false=%d  %.0f KB" % (
        name, K, D, int(out["done"].sum()), int((~out["connect"]).sum()), int((~out["connect_"]).sum()),
        os.path.getsize(path) / 1024))


def main():
    load_reference()
    # shipped behaviour: CoverageWorld() defaults (comm_r_scale 0.9, force off), 4 UAV / 20 PoI
    run_traj("ship_4x20_random", 4, 20, 300, "random", 0, compat=True)
    run_traj("ship_4x20_seek", 4, 20, 300, "seek", 1, compat=True)
    run_unit("ship_4x20_unit", 4, 20, 200, 2, compat=True)
    # BASELINE configs through the generalised make_world
    run_traj("gen_8x64_random", 8, 64, 150, "random", 3, obs_every=5)
    run_traj("gen_8x64_seek", 8, 64, 150, "seek", 4, obs_every=5)
    run_traj("gen_8x64_runaway", 8, 64, 60, "runaway", 5, obs_every=5)
    run_traj("gen_8x64_force", 8, 64, 150, "random", 6, comm_force_scale=1.0, obs_every=5)
    run_unit("gen_8x64_force_unit", 8, 64, 96, 7, comm_force_scale=1.0)
    run_traj("gen_16x256_force", 16, 256, 60, "random", 8, comm_force_scale=1.0, obs_every=20)
    run_unit("gen_16x256_force_unit", 16, 256, 12, 9, comm_force_scale=1.0, spread=1.6)
    # edges: BASELINE config 0 says 3 UAVs; N=2 (connect_ is always False); N=32 (one lane per UAV, max)
    run_traj("gen_3x20_force", 3, 20, 150, "random", 10, comm_force_scale=1.0, obs_every=3)
    run_traj("gen_2x5_force", 2, 5, 100, "random", 11, comm_force_scale=1.0)
    run_traj("gen_32x33_force", 32, 33, 30, "random", 12, comm_force_scale=1.0, obs_every=10)
    run_traj("gen_1x7", 1, 7, 60, "seek", 13)
    run_unit("gen_5x37_force_unit", 5, 37, 60, 14, comm_force_scale=0.5, comm_r_scale=0.8)
    # non-default cover / comm radii (dcc.yaml r_cover, r_comm)
    run_traj("gen_6x41_radii_force", 6, 41, 80, "random", 15, comm_r_scale=0.8, comm_force_scale=1.0, obs_every=4,
             r_cover=0.31, r_comm=0.27)
    run_unit("gen_7x23_radii_unit", 7, 23, 64, 16, comm_r_scale=0.6, comm_force_scale=0.3, r_cover=0.12, r_comm=0.55)
    # synthetic PoI layouts (uniform in [-1,1]^2, the reference's commented alternative, coverage.py:71): one layout per
    # file; the per-env-layout tests run several of these side by side in ONE vec-env
    for tag, seed in (("a", 21), ("b", 22), ("c", 23)):
        run_traj("layout_%s_5x12" % tag, 5, 12, 60, "seek" if tag == "b" else "random", seed, comm_force_scale=1.0,
                 obs_every=3, poi=np.random.default_rng(seed).uniform(-1, 1, (12, 2)))
    for tag, seed in (("a", 31), ("b", 32)):
        run_traj("layout_%s_8x64" % tag, 8, 64, 50, "seek" if tag == "b" else "random", seed, comm_force_scale=1.0,
                 obs_every=5, poi=np.random.default_rng(seed).uniform(-1, 1, (64, 2)))


if __name__ == "__main__":
    main()
