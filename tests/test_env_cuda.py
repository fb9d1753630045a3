"""GPU parity tests (run with -m gpu on the B200 box): the CUDA env path, called through the C ABI,
against (1) the golden vectors produced by the unmodified reference and (2) the CPU oracle on seeded
random batches.  Bar: observations, UAV/PoI state, done / connect / connect_ / adjacency bit-exact;
reward within 1e-6 relative of float32(reference float64 reward) (north_star asks 1e-5)."""
import numpy as np
import pytest
import torch

from golden_util import assert_step_matches, golden_cases, load_golden

pytestmark = pytest.mark.gpu

REW_RTOL = 1e-6


def _mk_cuda(g, E, numpy_compat=False):
    from dcc_b200.envs import CudaVecEnv
    c = g["cfg"]
    # golden cfg stores the WORLD's values; reference_compat=False passes ours through unchanged
    return CudaVecEnv(E, c["n_agents"], c["n_pois"], r_cover=c["r_cover"], r_comm=c["r_comm"],
                      comm_r_scale=c["comm_r_scale"], comm_force_scale=c["contact_force"] / 100.0,
                      reference_compat=False, pos_pois=g["poi"], numpy_compat=numpy_compat, want_connectivity=True)


def _result(env, obs, rew, done, infos):
    torch.cuda.synchronize()
    pv, en = env.get_state()
    cb = env.connect_bits.cpu().numpy()
    return dict(obs=obs.cpu().numpy(), reward=rew.cpu().numpy()[:, 0, 0], done=done.cpu().numpy()[:, 0],
                coverage_rate=infos.coverage_rate.cpu().numpy(), connect=(cb & 1).astype(bool),
                connect_=((cb >> 1) & 1).astype(bool), adj=env.adj.cpu().numpy().view(np.uint32),
                adj_=env.adj_s.cpu().numpy().view(np.uint32), pos_vel=pv, energy=en,
                rew_all=rew.cpu().numpy(), done_all=done.cpu().numpy())


@pytest.mark.parametrize("spec", [True, False])
@pytest.mark.parametrize("name", golden_cases("traj"))
def test_cuda_trajectory_vs_reference_golden(name, spec):
    g = load_golden(name)
    E = 3  # replicas of the same env: also checks env indexing
    env = _mk_cuda(g, E)
    if env.use_specialized(spec) != spec:
        pytest.skip("no specialised kernel for this shape (covered by the generic-kernel run)")
    obs0 = env.reset()
    torch.cuda.synchronize()
    for e in range(E):
        assert np.array_equal(obs0[e].cpu().numpy(), g["obs0"])
    obs_at = {int(t): k for k, t in enumerate(g["obs_steps"])}
    for t in range(g["cfg"]["T"]):
        a = torch.from_numpy(np.repeat(g["actions"][t][None], E, 0)).cuda()
        a_before = a.clone()
        r = _result(env, *env.step(a))
        assert torch.equal(a, a_before), "actions must not be mutated"
        k = obs_at.get(t)
        for e in range(E):
            assert_step_matches(name, r, g, t, e=e, obs_ref=None if k is None else g["obs"][k],
                                rew_rtol=REW_RTOL, rew_dtype=np.float32)
        assert np.all(r["rew_all"] == r["rew_all"][:, :1]) and np.all(r["done_all"] == r["done_all"][:, :1])
    env.close()


@pytest.mark.parametrize("name", golden_cases("unit"))
def test_cuda_unit_steps_vs_reference_golden(name):
    g = load_golden(name)
    K = g["cfg"]["K"]
    env = _mk_cuda(g, K)
    env.reset()
    env.set_state(g["pos_vel_in"], g["energy_in"])
    r = _result(env, *env.step(torch.from_numpy(g["actions"]).cuda()))
    for k in range(K):
        assert_step_matches(name, r, g, k, e=k, obs_ref=g["obs"][k], rew_rtol=REW_RTOL, rew_dtype=np.float32)
    env.close()


@pytest.mark.parametrize("N,M,E,force,steps,spec", [
    (8, 64, 4096, 0.0, 12, True), (8, 64, 2048, 1.0, 12, True), (8, 64, 2048, 1.0, 12, False),
    (16, 256, 512, 1.0, 6, True), (16, 256, 300, 1.0, 6, False), (4, 20, 4099, 0.0, 12, True), (4, 20, 4099, 1.0, 12, False),
    (3, 21, 257, 1.0, 8, False), (5, 9, 130, 1.0, 8, False), (32, 40, 64, 1.0, 4, False), (1, 3, 33, 0.0, 5, False),
    (2, 70, 65, 1.0, 6, False)])
def test_cuda_batch_vs_oracle(N, M, E, force, steps, spec):
    """Seeded random batch, several steps, vs the CPU oracle: everything bit-exact (reward: fp32 of the
    float64 oracle value up to 1e-6).  (3,21) and (5,9) have env blocks that are not 16-byte multiples
    and take the non-bulk store path; E values that are not multiples of the CTA size cover ragged grids."""
    from dcc_b200.envs import CudaVecEnv
    from oracle.env_oracle import OracleEnv
    rng = np.random.RandomState(N * 1000 + M)
    poi = rng.uniform(-1, 1, (M, 2))
    crs = 0.95
    env = CudaVecEnv(E, N, M, comm_r_scale=crs, comm_force_scale=force, reference_compat=False, pos_pois=poi,
                     want_connectivity=True)
    assert env.use_specialized(spec) == spec  # compile-time specialised kernel vs generic runtime-shape kernel
    orc = OracleEnv(E, N, M, poi, comm_r_scale=crs, contact_force=100.0 * force, n_threads=8)
    pv = np.zeros((E, N, 4)); pv[..., :2] = rng.uniform(-1.45, 1.45, (E, N, 2)) * rng.uniform(0.1, 1, (E, 1, 1))
    pv[..., 2:] = rng.uniform(-0.4, 0.4, (E, N, 2))
    en = rng.randint(0, 8, (E, M)).astype(np.uint8)
    en[::5] = np.where(rng.rand(*en[::5].shape) < 0.95, 6, 4)
    env.reset(); env.set_state(pv, en); orc.set_state(pv, en)
    n_done = 0
    for t in range(steps):
        a = (rng.standard_normal((E, N, 2)) * (1.0 + (t % 3))).astype(np.float32)
        r = _result(env, *env.step(torch.from_numpy(a).cuda()))
        o = orc.step(a)
        for key in ("done", "connect", "connect_", "adj", "adj_", "energy", "pos_vel", "obs"):
            assert np.array_equal(r[key], o[key]), "t=%d %s mismatch (%d envs)" % (
                t, key, int(np.sum(np.any((r[key] != o[key]).reshape(E, -1), axis=1))))
        ref = o["reward"].astype(np.float32)
        assert np.all(np.abs(r["reward"] - ref) <= REW_RTOL * np.maximum(1.0, np.abs(ref))), "t=%d reward" % t
        assert np.allclose(r["coverage_rate"], o["coverage_rate"], atol=1e-7)
        n_done += int(o["done"].sum())
    assert n_done > 0 or N <= 2
    env.close()


def test_cuda_numpy_compat_matches_reference_shapes_and_golden():
    g = load_golden("ship_4x20_random")
    env = _mk_cuda(g, 2, numpy_compat=True)
    obs = env.reset()
    assert obs.dtype == np.float64 and obs.shape == (2, 4, 110)
    assert np.array_equal(obs[0].astype(np.float32), g["obs0"])
    obs_at = {int(t): k for k, t in enumerate(g["obs_steps"])}
    for t in range(60):
        a = np.repeat(g["actions"][t][None], 2, 0)
        a0 = a.copy()
        obs, rew, done, infos = env.step(a)
        assert np.array_equal(a, a0)
        assert obs.dtype == np.float64 and rew.dtype == np.float64 and rew.shape == (2, 4, 1)
        assert done.dtype == bool and done.shape == (2, 4) and len(infos) == 2
        assert np.array_equal(obs[1].astype(np.float32), g["obs"][obs_at[t]])
        assert abs(rew[0, 0, 0] - np.float32(g["reward"][t])) <= REW_RTOL * max(1, abs(g["reward"][t]))
        assert bool(done[0, 0]) == bool(g["done"][t])
        assert abs(infos[1]["coverage_rate"] - g["coverage_rate"][t]) < 1e-6
    env.close()


def test_make_env_boundary():
    from argparse import Namespace
    from dcc_b200.envs import make_env
    cfg = Namespace(env_file="mpe.uav_dcc", env_class="DCEnv", scenario_name="coverage", num_agents=4, num_pois=20,
                    max_ep_len=150, r_cover=0.2, r_comm=0.4, comm_r_scale=0.95, comm_force_scale=0.0,
                    n_rollout_threads=16, seed=0)
    env = make_env(cfg)
    assert env.n_envs == 16 and env.n_agents == 4 and len(env.observation_space) == 4
    assert env.observation_space[0].shape == (110,) and env.share_observation_space[0].shape == (440,)
    assert env.action_space[0].__class__.__name__ == "Box" and env.action_space[0].shape == (2,)
    obs = env.reset()
    assert tuple(obs.shape) == (16, 4, 110)
    o, r, d, infos = env.step(torch.zeros(16, 4, 2, device="cuda"))
    assert tuple(r.shape) == (16, 4, 1) and tuple(d.shape) == (16, 4) and d.dtype == torch.bool
    assert "coverage_rate" in infos[0]
    env.close()
    cfg.per_env_layouts, cfg.poi_seed = True, 3          # optional key: one synthetic layout per env instance
    env = make_env(cfg)
    assert env.pos_pois_per_env.shape == (16, 20, 2)
    o = env.reset().cpu().numpy()
    q0 = o[:, 0, 4 + 2 * 3:4 + 2 * 3 + 2]                # first PoI offset of UAV 0 (at the origin) = that env's q_0
    assert np.array_equal(q0, env.pos_pois_per_env[:, 0].astype(np.float32)) and len(np.unique(q0[:, 0])) == 16
    env.close()
    cfg.env_file = "something_else"
    with pytest.raises(NotImplementedError):
        make_env(cfg)


def test_launch_geometry_does_not_change_results():
    from dcc_b200.envs import CudaVecEnv
    rng = np.random.RandomState(5)
    E, N, M = 1500, 8, 64
    poi = rng.uniform(-1, 1, (M, 2))
    a = torch.from_numpy(rng.standard_normal((4, E, N, 2)).astype(np.float32)).cuda()
    outs = []
    for wpc, ctas in ((4, 0), (1, 7), (8, 0), (16, 3), (2, 1000), (-4, 0), (-4, 5)):
        env = CudaVecEnv(E, N, M, comm_force_scale=1.0, reference_compat=False, pos_pois=poi)
        env.use_specialized(wpc < 0)
        env.set_launch(abs(wpc), ctas)
        env.reset()
        acc = []
        for t in range(4):
            o, r, d, i = env.step(a[t])
            acc += [o.clone(), r.clone(), d.clone(), i.coverage_rate.clone()]
        outs.append(acc)
        env.close()
    for other in outs[1:]:
        for x, y in zip(outs[0], other):
            assert torch.equal(x, y)


def _obs_from_state(pv, en, poi, N, M, m_energy=5.0):
    """Observation rows rebuilt from the compact state with torch on the device, in float64 then rounded once to
    float32 exactly as the reference stores them (scenarios/coverage.py:99-110 -> float32 buffer):
    [v_i, p_i, p_k - p_i (k != i), for every PoI: q_j - p_i, energy_j, m_energy, done_j]."""
    E = pv.shape[0]
    p, v = pv[..., :2], pv[..., 2:]
    rel = p[:, None, :, :] - p[:, :, None, :]                          # [e, i, k] = p_k - p_i
    keep = ~torch.eye(N, dtype=torch.bool, device=pv.device)
    rel = rel[:, keep].view(E, N, N - 1, 2).reshape(E, N, 2 * (N - 1))
    dq = poi[None, None] - p[:, :, None, :]                            # [e, i, j] = q_j - p_i
    enf = en.to(torch.float64)
    sec = torch.cat([dq, enf[:, None, :, None].expand(E, N, M, 1), torch.full((E, N, M, 1), m_energy, dtype=torch.float64, device=pv.device),
                     (enf >= m_energy).to(torch.float64)[:, None, :, None].expand(E, N, M, 1)], dim=-1).reshape(E, N, 5 * M)
    return torch.cat([v, p, rel, sec], dim=-1).to(torch.float32)


@pytest.mark.parametrize("N,M,E,force", [(8, 64, 65536, 0.0), (16, 256, 32768, 1.0)])
def test_full_size_properties(N, M, E, force):
    """BASELINE.json's full sizes (configs[1] and [2]), where the CPU oracle would take minutes: size-independent
    properties of the step instead — (1) every observation row is exactly the function of the compact state the
    reference defines, (2) reward / done are shared by the agents of an env, (3) a done env comes back reset,
    (4) PoI energy never decreases inside an episode and the coverage rate is the done-PoI fraction, (5) the first
    4096 envs reproduce a 4096-env run bit for bit (what sharding the env axis across GPUs relies on), and that
    small run is itself checked against the CPU oracle, (6) connect bits agree with a float64 union-find-free
    recomputation (adjacency powers) on a sample."""
    from dcc_b200.envs import CudaVecEnv
    from oracle.env_oracle import OracleEnv
    rng = np.random.RandomState(7)
    poi = rng.uniform(-1, 1, (M, 2))
    Es = 4096
    kw = dict(comm_r_scale=0.95, comm_force_scale=force, reference_compat=False, pos_pois=poi, want_connectivity=True)
    env, small = CudaVecEnv(E, N, M, **kw), CudaVecEnv(Es, N, M, **kw)
    orc = OracleEnv(Es, N, M, poi, comm_r_scale=0.95, contact_force=100.0 * force, n_threads=8)
    dev = env.device
    poi_t = torch.from_numpy(poi).to(dev)
    g = torch.Generator(device=dev).manual_seed(3)
    env.reset(); small.reset(); orc.reset()
    # start from a spread-out state (positions up to the |p| > 1.5 bound, PoIs close to completion in every 5th env)
    # so that both episode-end conditions fire within the 40 steps
    pv0 = np.zeros((E, N, 4)); pv0[..., :2] = rng.uniform(-1.45, 1.45, (E, N, 2)) * rng.uniform(0.1, 1, (E, 1, 1))
    pv0[..., 2:] = rng.uniform(-0.4, 0.4, (E, N, 2))
    en0 = rng.randint(0, 5, (E, M)).astype(np.uint8)
    en0[::5] = np.where(rng.rand(*en0[::5].shape) < 0.97, 6, 4)
    env.set_state(pv0, en0); small.set_state(pv0[:Es], en0[:Es]); orc.set_state(pv0[:Es], en0[:Es])
    ptrs = lambda e: e.get_state()    # noqa: E731
    prev_en = torch.from_numpy(en0).to(dev)
    n_done = 0
    for t in range(40):
        a = torch.randn((E, N, 2), generator=g, device=dev) * (2.0 if t % 7 == 6 else 1.0)
        obs, rew, done, infos = env.step(a)
        pv_h, en_h = ptrs(env)
        pv, en = torch.from_numpy(pv_h).to(dev), torch.from_numpy(en_h).to(dev)
        assert torch.equal(obs, _obs_from_state(pv, en, poi_t, N, M)), "t=%d obs is not f(state)" % t      # (1)
        assert bool((rew == rew[:, :1]).all()) and bool((done == done[:, :1]).all())                       # (2)
        d = done[:, 0]
        n_done += int(d.sum())
        if bool(d.any()):                                                                                  # (3)
            assert bool((pv[d] == 0).all()) and bool((en[d] == 0).all())
        assert bool((en[~d] >= prev_en[~d]).all())                                                         # (4)
        live_cov = (en >= 5).float().mean(1)
        assert torch.allclose(infos.coverage_rate[~d], live_cov[~d], atol=1e-6)
        prev_en = en
        o2, r2, d2, i2 = small.step(a[:Es].contiguous())                                                   # (5)
        assert torch.equal(o2, obs[:Es]) and torch.equal(r2, rew[:Es]) and torch.equal(d2, done[:Es])
        assert torch.equal(small.connect_bits, env.connect_bits[:Es]) and torch.equal(small.adj, env.adj[:Es])
        if t % 8 == 0 or t == 39:
            o = orc.step(a[:Es].cpu().numpy())
            assert np.array_equal(o2.cpu().numpy(), o["obs"]) and np.array_equal(d2.cpu().numpy()[:, 0], o["done"])
            cb = small.connect_bits.cpu().numpy()
            assert np.array_equal((cb & 1).astype(bool), o["connect"])                                     # (6)
            assert np.array_equal(((cb >> 1) & 1).astype(bool), o["connect_"])
            ref = o["reward"].astype(np.float32)
            assert np.all(np.abs(r2.cpu().numpy()[:, 0, 0] - ref) <= REW_RTOL * np.maximum(1.0, np.abs(ref)))
        else:
            orc.step(a[:Es].cpu().numpy())
    assert n_done > 0
    env.close(); small.close()


def test_step_host_equals_device_step():
    """The host-buffer entry point (what the e2e bench and numpy_compat use) returns exactly the device path's results."""
    from dcc_b200.envs import CudaVecEnv
    rng = np.random.RandomState(11)
    E, N, M = 3000, 8, 64
    poi = rng.uniform(-1, 1, (M, 2))
    a_env, b_env = (CudaVecEnv(E, N, M, comm_force_scale=1.0, reference_compat=False, pos_pois=poi) for _ in range(2))
    a_env.reset(); b_env.reset()
    for t in range(6):
        a = (rng.standard_normal((E, N, 2)) * 1.5).astype(np.float32)
        obs, rew, done, infos = a_env.step(torch.from_numpy(a).cuda())
        ho, hr, hd, hc = b_env.step_host(a)
        assert np.array_equal(obs.cpu().numpy(), ho) and np.array_equal(rew.cpu().numpy()[:, :, 0], hr)
        assert np.array_equal(done.cpu().numpy(), hd.astype(bool)) and np.array_equal(infos.coverage_rate.cpu().numpy(), hc)
    a_env.close(); b_env.close()


def test_render_snapshot_and_connectivity_outputs():
    """f-3: headless render / snapshot of the compact state; adjacency outputs can be switched on after creation."""
    from dcc_b200.envs import CudaVecEnv
    from oracle.env_oracle import OracleEnv
    rng = np.random.RandomState(2)
    E, N, M = 5, 4, 20
    poi = rng.uniform(-1, 1, (M, 2))
    env = CudaVecEnv(E, N, M, comm_force_scale=0.0, reference_compat=False, pos_pois=poi)
    orc = OracleEnv(E, N, M, poi, comm_r_scale=0.95, contact_force=0.0, n_threads=1)
    env.reset(); orc.reset()
    assert env.render("human") is None
    f0 = env.render("rgb_array")
    assert len(f0) == 1 and f0[0][0].shape == (350, 350, 3) and f0[0][0].dtype == np.uint8
    env.enable_connectivity_outputs()
    for t in range(12):
        a = (rng.standard_normal((E, N, 2)) * 2).astype(np.float32)
        env.step(torch.from_numpy(a).cuda())
        o = orc.step(a)
    st = env.snapshot()
    assert np.array_equal(st["pos_vel"], orc.pos_vel) and np.array_equal(st["energy"], orc.energy)
    assert np.array_equal(st["adj"], o["adj"]) and np.array_equal((st["connect_bits"] & 1).astype(bool), o["connect"])
    frames = env.render("rgb_array", max_envs=3, size=128)
    assert len(frames) == 3 and frames[2][0].shape == (128, 128, 3) and (frames[0][0] != f0[0][0][:128, :128]).any()
    env.close()


def test_cuda_random_shapes_and_parameters_vs_oracle():
    """Seeded sweep over shapes the fixed list above does not hit (N up to 32, M up to 300, prime / ragged sizes) AND over
    the env parameters (cover / comm radii, comm_r_scale, pull force), several steps each from a spread-out state: the
    runtime-shape kernel against the CPU oracle, everything bit-exact."""
    from dcc_b200.envs import CudaVecEnv
    from oracle.env_oracle import OracleEnv
    master = np.random.RandomState(2024)
    for case in range(14):
        N = int(master.choice([1, 2, 3, 5, 6, 7, 9, 11, 13, 17, 24, 31, 32]))
        M = int(master.choice([1, 2, 4, 17, 31, 33, 50, 97, 128, 199, 300]))
        E = int(master.choice([1, 2, 31, 33, 100, 257]))
        r_cover, r_comm = float(master.uniform(0.05, 0.5)), float(master.uniform(0.1, 0.8))
        crs, force = float(master.uniform(0.5, 1.0)), float(master.choice([0.0, 0.3, 1.0, 2.5]))
        rng = np.random.RandomState(1000 + case)
        poi = rng.uniform(-1, 1, (M, 2))
        env = CudaVecEnv(E, N, M, r_cover=r_cover, r_comm=r_comm, comm_r_scale=crs, comm_force_scale=force,
                         reference_compat=False, pos_pois=poi, want_connectivity=True)
        orc = OracleEnv(E, N, M, poi, r_cover=r_cover, r_comm=r_comm, comm_r_scale=crs, contact_force=100.0 * force, n_threads=4)
        pv = np.zeros((E, N, 4)); pv[..., :2] = rng.uniform(-1.4, 1.4, (E, N, 2)) * rng.uniform(0.05, 1, (E, 1, 1))
        pv[..., 2:] = rng.uniform(-0.4, 0.4, (E, N, 2))
        en = rng.randint(0, 7, (E, M)).astype(np.uint8)
        env.reset(); env.set_state(pv, en); orc.set_state(pv, en)
        tag = "case %d N=%d M=%d E=%d rc=%.3f rm=%.3f crs=%.3f f=%.1f" % (case, N, M, E, r_cover, r_comm, crs, force)
        for t in range(6):
            a = (rng.standard_normal((E, N, 2)) * (0.5 + t % 3)).astype(np.float32)
            r = _result(env, *env.step(torch.from_numpy(a).cuda()))
            o = orc.step(a)
            for key in ("done", "connect", "connect_", "adj", "adj_", "energy", "pos_vel", "obs"):
                assert np.array_equal(r[key], o[key]), "%s t=%d %s mismatch" % (tag, t, key)
            ref = o["reward"].astype(np.float32)
            assert np.all(np.abs(r["reward"] - ref) <= REW_RTOL * np.maximum(1.0, np.abs(ref))), "%s t=%d reward" % (tag, t)
        env.close()


@pytest.mark.parametrize("shape,tags,spec", [("5x12", "abc", False), ("8x64", "ab", True), ("8x64", "ab", False)])
def test_cuda_per_env_poi_layouts_vs_reference_golden(shape, tags, spec):
    """dcc_env_set_poi_layouts: one vec-env whose instances have DIFFERENT PoI layouts; instance e replays the golden
    the unmodified reference recorded on layout e (several replicas of each, interleaved), then the env is switched
    back to the shared layout."""
    gs = [load_golden("layout_%s_%s" % (t, shape)) for t in tags]
    c = gs[0]["cfg"]
    reps = 3
    E, T = len(gs) * reps, min(g["cfg"]["T"] for g in gs)
    which = [e % len(gs) for e in range(E)]                  # env e runs golden which[e]
    env = _mk_cuda(gs[0], E)
    assert env.use_specialized(spec) == spec
    env.set_poi_layouts(np.stack([gs[w]["poi"] for w in which]))
    obs0 = env.reset().cpu().numpy()
    for e in range(E):
        assert np.array_equal(obs0[e], gs[which[e]]["obs0"])
    for t in range(T):
        a = np.stack([gs[w]["actions"][t] for w in which])
        r = _result(env, *env.step(torch.from_numpy(a).cuda()))
        for e in range(E):
            g = gs[which[e]]
            obs_at = {int(tt): k for k, tt in enumerate(g["obs_steps"])}
            k = obs_at.get(t)
            assert_step_matches("layout %s env %d" % (shape, e), r, g, t, e=e, obs_ref=None if k is None else g["obs"][k],
                                rew_rtol=REW_RTOL, rew_dtype=np.float32)
    # back to the shared layout (the one given at creation = golden 0's)
    env.set_poi_layouts(None)
    obs0 = env.reset().cpu().numpy()
    assert all(np.array_equal(obs0[e], gs[0]["obs0"]) for e in range(E))
    with pytest.raises(ValueError):
        env.set_poi_layouts(np.zeros((E + 1, c["n_pois"], 2)))
    env.close()


@pytest.mark.parametrize("name", ["gen_8x64_runaway", "ship_4x20_seek", "gen_8x64_force"])
def test_rollout_insert_vs_reference_golden(name):
    """Row a13 (learner.py:254-276 + shared_buffer.py:72-105): what `Learner.insert` stores per step — the shared reward
    and mask = 1 - done — asserted against the reference trajectory's own reward / done records (runaway = episodes
    that end, so masks of 0 appear), through SharedReplayBuffer.insert_env_step (dcc_rollout_insert) on the buffer
    slices the rollout uses; obs[t+1] is the env kernel's in-place output (post-auto-reset observation)."""
    from argparse import Namespace
    from dcc_b200.buffer import SharedReplayBuffer
    from dcc_b200.envs.spaces import Box
    g = load_golden(name)
    c = g["cfg"]
    E, T, N = 3, c["T"], c["n_agents"]
    env = _mk_cuda(g, E)
    D = env.obs_dim
    cfg = Namespace(max_ep_len=T, n_rollout_threads=E, gamma=0.99, gae_lambda=0.95, num_agents=N, device=0)
    buf = SharedReplayBuffer(cfg, Box(-np.inf, np.inf, (D,)), Box(-np.inf, np.inf, (N * D,)), Box(-1, 1, (2,)))
    env.reset(out_obs=buf.obs[0])
    for t in range(T):
        a = torch.from_numpy(np.repeat(g["actions"][t][None], E, 0)).cuda()
        assert buf.step == t
        obs, rew, done, infos = env.step(a, out_obs=buf.obs[t + 1])
        buf.insert_env_step(rew, done.view(torch.uint8))
    torch.cuda.synchronize()
    assert buf.step == 0                                                    # wrapped after T inserts
    rew = buf.rewards_te.cpu().numpy()
    masks = buf.masks_te.cpu().numpy()
    ref_rew = g["reward"].astype(np.float32)[:T]
    ref_done = g["done"].astype(bool)[:T]
    assert masks.shape == (T + 1, E) and np.all(masks[0] == 1.0)
    for e in range(E):
        assert np.all(np.abs(rew[:, e] - ref_rew) <= REW_RTOL * np.maximum(1.0, np.abs(ref_rew))), name
        assert np.array_equal(masks[1:, e], 1.0 - ref_done.astype(np.float32)), name
    if "runaway" in name:
        assert ref_done.any()
    # the reference-shaped views expose the same numbers per agent: (T, E, N, 1)
    assert buf.rewards.shape == (T, E, N, 1) and torch.equal(buf.rewards[:, :, 0, 0], buf.rewards_te)
    assert buf.masks.shape == (T + 1, E, N, 1) and torch.equal(buf.masks[:, :, N - 1, 0], buf.masks_te)
    # obs[t+1] holds the golden's post-reset observations where recorded
    obs_at = {int(t): k for k, t in enumerate(g["obs_steps"])}
    for t, k in obs_at.items():
        if t < T:
            assert np.array_equal(buf.obs[t + 1, 1].cpu().numpy(), g["obs"][k])
    env.close()
