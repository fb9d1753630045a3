"""Shared by the MAPPO golden generator (reference side) and the parity tests (product side):
deterministic parameter recipe with the reference's state_dict key names."""
import numpy as np


def _maybe_drop_feature_norm(shapes, feature_norm, layer_N=1, hidden=None, recurrent_N=0):
    """Drops the feature_norm entries (use_feature_normalization: false), inserts the fc2.{1..layer_N-1} blocks
    after fc2.0 (layer_N > 1) and, for recurrent policies (recurrent_N > 0), the RNNLayer's entries between the trunk and
    the head (rnn.py:13-22: torch.nn.GRU's weight_ih / weight_hh / bias_ih / bias_hh per layer, then the LayerNorm),
    keeping the reference's state_dict order."""
    out = {}
    for k, v in shapes.items():
        if not feature_norm and k.startswith("base.feature_norm"):
            continue
        out[k] = v
        if k == "base.mlp.fc2.0.2.bias":
            for i in range(1, layer_N):
                out["base.mlp.fc2.%d.0.weight" % i] = (hidden, hidden)
                out["base.mlp.fc2.%d.0.bias" % i] = (hidden,)
                out["base.mlp.fc2.%d.2.weight" % i] = (hidden,)
                out["base.mlp.fc2.%d.2.bias" % i] = (hidden,)
            for l in range(recurrent_N):
                out["rnn.rnn.weight_ih_l%d" % l] = (3 * hidden, hidden)
                out["rnn.rnn.weight_hh_l%d" % l] = (3 * hidden, hidden)
                out["rnn.rnn.bias_ih_l%d" % l] = (3 * hidden,)
                out["rnn.rnn.bias_hh_l%d" % l] = (3 * hidden,)
            if recurrent_N:
                out["rnn.norm.weight"] = (hidden,)
                out["rnn.norm.bias"] = (hidden,)
    return out


def actor_param_shapes(obs_dim, hidden, act_dim=2, feature_norm=True, layer_N=1, recurrent_N=0):
    return _maybe_drop_feature_norm({
        "base.feature_norm.weight": (obs_dim,), "base.feature_norm.bias": (obs_dim,),
        "base.mlp.fc1.0.weight": (hidden, obs_dim), "base.mlp.fc1.0.bias": (hidden,),
        "base.mlp.fc1.2.weight": (hidden,), "base.mlp.fc1.2.bias": (hidden,),
        "base.mlp.fc2.0.0.weight": (hidden, hidden), "base.mlp.fc2.0.0.bias": (hidden,),
        "base.mlp.fc2.0.2.weight": (hidden,), "base.mlp.fc2.0.2.bias": (hidden,),
        "act.action_out.fc_mean.weight": (act_dim, hidden), "act.action_out.fc_mean.bias": (act_dim,),
        "act.action_out.logstd._bias": (act_dim, 1),
    }, feature_norm, layer_N, hidden, recurrent_N)


def critic_param_shapes(share_dim, hidden, feature_norm=True, layer_N=1, recurrent_N=0):
    return _maybe_drop_feature_norm({
        "base.feature_norm.weight": (share_dim,), "base.feature_norm.bias": (share_dim,),
        "base.mlp.fc1.0.weight": (hidden, share_dim), "base.mlp.fc1.0.bias": (hidden,),
        "base.mlp.fc1.2.weight": (hidden,), "base.mlp.fc1.2.bias": (hidden,),
        "base.mlp.fc2.0.0.weight": (hidden, hidden), "base.mlp.fc2.0.0.bias": (hidden,),
        "base.mlp.fc2.0.2.weight": (hidden,), "base.mlp.fc2.0.2.bias": (hidden,),
        "v_out.weight": (1, hidden), "v_out.bias": (1,),
    }, feature_norm, layer_N, hidden, recurrent_N)


def net_shapes(c):
    """(actor shapes, critic shapes) for a golden's cfg dict: honours use_feature_normalization and use_centralized_V
    (a decentralised critic reads one agent's observation, learner.py:43-46)."""
    fn = c.get("use_feature_normalization", True)
    share = c["n_agents"] * c["obs_dim"] if c.get("use_centralized_V", True) else c["obs_dim"]
    ln = c.get("layer_N", 1)
    rn = int(c.get("recurrent_N", 1)) if (c.get("use_recurrent_policy", False) or c.get("use_naive_recurrent_policy", False)) else 0
    return (actor_param_shapes(c["obs_dim"], c["hidden"], feature_norm=fn, layer_N=ln, recurrent_N=rn),
            critic_param_shapes(share, c["hidden"], feature_norm=fn, layer_N=ln, recurrent_N=rn))


def make_params(shapes, seed):
    """Seeded, platform-independent parameter values (numpy Generator): non-trivial LN affine, biases and logstd so
    that every term of the forward/backward is exercised (the reference's own init has zero biases and logstd)."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shp in shapes.items():
        if name.endswith("logstd._bias"):
            v = rng.normal(0.0, 0.15, shp)
        elif "feature_norm.weight" in name or name.endswith(".2.weight") or name == "rnn.norm.weight":   # LayerNorm gains
            v = 1.0 + rng.normal(0.0, 0.1, shp)
        elif name.startswith("rnn.rnn.bias"):
            v = rng.normal(0.0, 0.05, shp)
        elif name.startswith("rnn.rnn.weight"):     # GRU matrices (3H, H)
            v = rng.normal(0.0, 1.0 / np.sqrt(shp[1]), shp)
        elif name.endswith("bias"):
            v = rng.normal(0.0, 0.05, shp)
        elif "fc_mean.weight" in name:
            v = rng.normal(0.0, 0.3 / np.sqrt(shp[1]), shp)
        else:  # Linear weights (out, in)
            v = rng.normal(0.0, np.sqrt(2.0 / shp[1]), shp)
        out[name] = v.astype(np.float32)
    return out


def sample_tensor(a, max_full=4096):
    """Golden-side compression of a big 'after' tensor: strided sample + float64 sum / sum of squares."""
    flat = np.asarray(a, dtype=np.float32).reshape(-1)
    stride = 1 if flat.size <= max_full else int(np.ceil(flat.size / max_full))
    return dict(stride=stride, sample=flat[::stride].copy(), sum=float(flat.astype(np.float64).sum()),
                sumsq=float((flat.astype(np.float64) ** 2).sum()))
