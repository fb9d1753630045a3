"""CPU tests of the compact-state algebra restatement (oracle/compact_oracle.py) that the CUDA compact path
(csrc/dcc_compact.cuh) is checked against: the observation rows rebuilt from the compact state are bit-identical to
the UNMODIFIED reference's recorded observations (env goldens), and the folded first layer / unfolded weight gradient
equal the direct ones (float64 identities; the only difference is float32 rounding of the stored observation values)."""
import json
import os

import numpy as np
import pytest

from golden_util import golden_cases, load_golden
from oracle import compact_oracle as co

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", golden_cases("traj"))
def test_obs_rows_from_state_equal_reference_observations(name):
    g = load_golden(name)
    c = g["cfg"]
    for k, t in enumerate(g["obs_steps"]):
        obs = co.obs_rows(g["pos_vel"][int(t)][None], g["energy"][int(t)][None], g["poi"])[0]
        assert np.array_equal(obs, g["obs"][k].astype(np.float32)), (name, int(t))
    obs0 = co.obs_rows(np.zeros((1, c["n_agents"], 4)), np.zeros((1, c["n_pois"])), g["poi"])[0]
    assert np.array_equal(obs0, g["obs0"].astype(np.float32))


@pytest.mark.parametrize("N,M", [(8, 64), (4, 20), (3, 7), (1, 5), (16, 33)])
@pytest.mark.parametrize("centralized", [False, True])
@pytest.mark.parametrize("normalize", [True, False])
def test_fold_and_unfold_identities(N, M, centralized, normalize):
    rng = np.random.default_rng(N * 100 + M)
    R, H = 6, 10
    poi = rng.uniform(-1, 1, (M, 2))
    pv = rng.normal(0, 0.6, (R, N, 4))
    en = rng.integers(0, 9, (R, M))
    x = co.obs_rows(pv, en, poi).astype(np.float64)
    D = x.shape[2]
    nb = N if centralized else 1
    rows = x.reshape(R, N * D) if centralized else x.reshape(R * N, D)
    xh = (rows - rows.mean(1, keepdims=True)) / np.sqrt(rows.var(1, keepdims=True) + co.LN_EPS) if normalize else rows
    f = co.features(pv, en, poi, centralized, normalize)
    own, K = co.feature_dims(N, M, nb)
    assert f.shape == (rows.shape[0], K) and K < rows.shape[1]
    Wg = rng.normal(0, 0.1, (H, rows.shape[1]))
    Wt = co.fold_weights(Wg, poi, N, M, nb)
    assert np.abs(xh @ Wg.T - f @ Wt.T).max() < 2e-6       # float32 rounding of q_j - p_i in the stored observations
    dz = rng.normal(0, 1, (rows.shape[0], H))
    assert np.abs(dz.T @ xh - co.unfold_grad(dz.T @ f, poi, N, M, nb)).max() < 5e-6


def test_state_from_recorded_observations_round_trips():
    z = np.load(os.path.join(GOLDEN, "mappo_gen_8x64_h256.npz"))
    c = json.loads(str(z["cfg"]))
    obs = z["it1_obs"]
    N, M = c["n_agents"], c["n_pois"]
    pv, en = co.state_from_obs(obs, N, M)
    from dcc_b200.envs.cuda_vec_env import reference_pois
    back = co.obs_rows(pv.reshape(-1, N, 4), en.reshape(-1, M), reference_pois(M)).reshape(obs.shape)
    assert np.abs(back - obs).max() <= 1.2e-7          # positions known to float32 only: relative offsets move by <= 1 ulp
    own = 2 * N + 2
    assert np.array_equal(back[..., :4], obs[..., :4]) and np.array_equal(back[..., own + 2::5], obs[..., own + 2::5])
