"""Pins the CPU oracle (oracle/dcc_env_oracle.c) to the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  Observations, UAV state, PoI energies and every flag must be bit-exact;
the reward to 1e-12 relative (summation order only)."""
import numpy as np
import pytest

from golden_util import assert_step_matches, golden_cases, load_golden
from oracle.env_oracle import OracleEnv


def _mk(g, E=1, n_threads=1):
    c = g["cfg"]
    return OracleEnv(E, c["n_agents"], c["n_pois"], g["poi"], c["r_cover"], c["r_comm"], c["comm_r_scale"],
                     c["contact_force"], n_threads=n_threads)


@pytest.mark.parametrize("name", golden_cases("traj"))
def test_oracle_trajectory(name):
    g = load_golden(name)
    env = _mk(g)
    obs0 = env.reset()
    assert np.array_equal(obs0[0], g["obs0"])
    obs_at = {int(t): k for k, t in enumerate(g["obs_steps"])}
    for t in range(g["cfg"]["T"]):
        r = env.step(g["actions"][t][None])
        k = obs_at.get(t)
        assert_step_matches(name, r, g, t, obs_ref=None if k is None else g["obs"][k])


@pytest.mark.parametrize("name", golden_cases("unit"))
def test_oracle_unit_steps(name):
    g = load_golden(name)
    K = g["cfg"]["K"]
    env = _mk(g, E=K, n_threads=3)          # all K injected states at once, also covers the threaded path
    env.set_state(g["pos_vel_in"], g["energy_in"])
    r = env.step(g["actions"])
    for k in range(K):
        assert_step_matches(name, r, g, k, e=k, obs_ref=g["obs"][k])


def test_oracle_actions_not_mutated_and_threads_agree():
    g = load_golden("gen_8x64_force")
    rng = np.random.RandomState(0)
    E = 37
    a = rng.standard_normal((E, 8, 2)).astype(np.float32)
    a0 = a.copy()
    e1, e2 = _mk(g, E, 1), _mk(g, E, 4)
    pv = rng.uniform(-1.2, 1.2, (E, 8, 4)); en = rng.randint(0, 8, (E, 64))
    e1.set_state(pv, en); e2.set_state(pv, en)
    r1, r2 = e1.step(a), e2.step(a)
    assert np.array_equal(a, a0)
    for k in ("obs", "reward", "done", "pos_vel", "energy", "connect_bits", "adj", "adj_"):
        assert np.array_equal(r1[k], r2[k]), k


@pytest.mark.parametrize("shape,tags", [("5x12", "abc"), ("8x64", "ab")])
def test_oracle_per_env_poi_layouts(shape, tags):
    """One vec-env whose instances have DIFFERENT PoI layouts: env e replays the golden recorded by the unmodified
    reference on layout e (tests/golden/env_layout_*), all stepped together."""
    gs = [load_golden("layout_%s_%s" % (t, shape)) for t in tags]
    c = gs[0]["cfg"]
    E, T = len(gs), min(g["cfg"]["T"] for g in gs)
    poi = np.stack([g["poi"] for g in gs])
    orc = OracleEnv(E, c["n_agents"], c["n_pois"], poi, c["r_cover"], c["r_comm"], c["comm_r_scale"], c["contact_force"])
    obs0 = orc.reset()
    for e, g in enumerate(gs):
        assert np.array_equal(obs0[e], g["obs0"])
    for t in range(T):
        r = orc.step(np.stack([g["actions"][t] for g in gs]))
        for e, g in enumerate(gs):
            obs_at = {int(tt): k for k, tt in enumerate(g["obs_steps"])}
            k = obs_at.get(t)
            assert_step_matches("layout %s env %d" % (shape, e), r, g, t, e=e, obs_ref=None if k is None else g["obs"][k])
