"""Pins the NumPy MAPPO oracle (oracle/mappo_oracle.py) to golden vectors recorded from the unmodified
reference learner (tests/golden/make_golden_mappo.py)."""
import glob
import json
import os

import numpy as np
import pytest

from mappo_util import actor_param_shapes, critic_param_shapes, make_params
from oracle import mappo_oracle as mo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[6:-4] for p in glob.glob(os.path.join(GOLDEN, "mappo_*.npz")))


def load(name):
    z = np.load(os.path.join(GOLDEN, "mappo_%s.npz" % name))
    g = {k: z[k] for k in z.files}
    g["cfg"] = json.loads(str(g["cfg"]))
    return g


def check_params(tag, params, g, prefix, rtol=2e-5, atol=2e-6):
    for k, v in params.items():
        key = prefix + k
        if key + ":sample" not in g:
            continue
        stride, s, ss = g[key + ":meta"]
        flat = np.asarray(v, dtype=np.float64).reshape(-1)
        ref = g[key + ":sample"].astype(np.float64)
        got = flat[::int(stride)]
        assert np.allclose(got, ref, rtol=rtol, atol=atol), "%s %s max|d|=%g" % (tag, k, np.abs(got - ref).max())
        assert abs(flat.sum() - s) <= 1e-4 * max(1.0, np.abs(flat).sum()), (tag, k, "sum")
        assert abs((flat ** 2).sum() - ss) <= 1e-4 * max(1.0, ss), (tag, k, "sumsq")


@pytest.mark.parametrize("name", CASES)
def test_mappo_oracle_vs_reference(name):
    g = load(name)
    c = g["cfg"]
    N, D, Hd = c["n_agents"], c["obs_dim"], c["hidden"]
    ap = make_params(actor_param_shapes(D, Hd), c["actor_seed"])
    cp = make_params(critic_param_shapes(N * D, Hd), c["critic_seed"])
    tr = mo.Trainer(ap, cp, c)
    for it in range(1, c["iters"] + 1):
        p = "it%d_" % it
        obs, act = g[p + "obs"], g[p + "actions"]
        T, E = act.shape[:2]
        assert np.allclose(tr.vn.state(), g[p + "vn_before"], rtol=1e-5, atol=1e-12)
        # rollout forward (teacher-forced on the recorded actions): log-probs and values
        mean = tr.actor.forward(obs[:-1].reshape(T * E * N, D))
        logp, _ = mo.gaussian_logp_entropy(mean, tr.actor.p["act.action_out.logstd._bias"].reshape(1, -1),
                                           act.reshape(-1, 2))
        assert np.allclose(logp.reshape(T, E, N, 1), g[p + "logp"], rtol=1e-5, atol=2e-5)
        v = tr.critic.forward(obs.reshape((T + 1) * E, N * D)).reshape(T + 1, E, 1, 1)
        assert np.allclose(np.broadcast_to(v, g[p + "value_preds"].shape), g[p + "value_preds"], rtol=1e-5, atol=2e-5)
        # GAE
        ret = mo.gae_returns(g[p + "rewards"], g[p + "value_preds"], g[p + "masks"], tr.vn, c["gamma"], c["gae_lambda"])
        assert np.allclose(ret[:-1], g[p + "returns"][:-1], rtol=1e-5, atol=1e-4)
        # update
        info = tr.train(obs, act, g[p + "logp"], g[p + "value_preds"], g[p + "returns"], float(g[p + "lr"]),
                        c["ppo_epoch"])
        ref = dict(zip(("value_loss", "policy_loss", "dist_entropy", "actor_grad_norm", "critic_grad_norm", "ratio"),
                       g[p + "train_info"]))
        for k in ref:
            assert abs(info[k] - ref[k]) <= 3e-5 * max(1.0, abs(ref[k])), (it, k, info[k], ref[k])
        assert np.allclose(tr.vn.state(), g[p + "vn_after"], rtol=1e-5, atol=1e-12)
        check_params("actor it%d" % it, tr.actor.p, g, p + "actor.")
        check_params("critic it%d" % it, tr.critic.p, g, p + "critic.")
