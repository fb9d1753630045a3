"""Helpers shared by the oracle (CPU) and CUDA (GPU) parity tests: load tests/golden/env_*.npz and
replay a golden case through any backend exposing set_state/reset/step with the OracleEnv result keys."""
import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases(kind=None):
    out = []
    for p in sorted(glob.glob(os.path.join(GOLDEN_DIR, "env_*.npz"))):
        name = os.path.basename(p)[4:-4]
        is_unit = name.endswith("_unit")
        if kind is None or (kind == "unit") == is_unit:
            out.append(name)
    return out


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, "env_%s.npz" % name))
    g = {k: z[k] for k in z.files}
    g["cfg"] = json.loads(str(g["cfg"]))
    return g


def assert_step_matches(tag, r, g, t, e=0, obs_ref=None, rew_rtol=1e-12, rew_dtype=np.float64):
    """Compare one backend step result `r` (env index e) with golden record t.
    Flags / integer state / observations: bit-exact.  Reward: relative tolerance."""
    assert bool(r["done"][e]) == bool(g["done"][t]), "%s t=%d done" % (tag, t)
    assert bool(r["connect"][e]) == bool(g["connect"][t]), "%s t=%d connect" % (tag, t)
    assert bool(r["connect_"][e]) == bool(g["connect_"][t]), "%s t=%d connect_" % (tag, t)
    if r.get("adj") is not None:
        assert np.array_equal(r["adj"][e], g["adj"][t]), "%s t=%d adj" % (tag, t)
        assert np.array_equal(r["adj_"][e], g["adj_"][t]), "%s t=%d adj_" % (tag, t)
    assert np.array_equal(r["energy"][e], g["energy"][t]), "%s t=%d energy" % (tag, t)
    if r.get("energy_pre") is not None:
        assert np.array_equal(r["energy_pre"][e], g["energy_pre"][t]), "%s t=%d energy_pre" % (tag, t)
        assert np.array_equal(r["pos_vel_pre"][e], g["pos_vel_pre"][t]), "%s t=%d pos_vel_pre" % (tag, t)
    assert np.array_equal(r["pos_vel"][e], g["pos_vel"][t]), "%s t=%d pos_vel max|d|=%g" % (
        tag, t, np.abs(r["pos_vel"][e] - g["pos_vel"][t]).max())
    ref = rew_dtype(g["reward"][t])
    got = r["reward"][e]
    assert abs(float(got) - float(ref)) <= rew_rtol * max(1.0, abs(float(ref))), "%s t=%d reward %r vs %r" % (
        tag, t, got, ref)
    cov = float(r["coverage_rate"][e])
    assert abs(cov - float(g["coverage_rate"][t])) <= 1e-7, "%s t=%d coverage_rate" % (tag, t)
    if obs_ref is not None:
        assert np.array_equal(r["obs"][e], obs_ref), "%s t=%d obs max|d|=%g" % (
            tag, t, np.abs(r["obs"][e].astype(np.float64) - obs_ref).max())
