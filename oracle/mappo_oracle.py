"""NumPy restatement of the reference's MAPPO math.  TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/ (and bench.py's CPU-baseline leg) import this module; the product package never does.
It restates, in float64 NumPy with hand-written backward passes (paths relative to
/root/reference/uav_dcc_control/):

  actor / critic forward      algos/r_actor_critic.py:43-57,111-121 -> algos/algo_utils/mlp.py:25-29,52-58
                              -> algos/algo_utils/act.py:79-84,165-184 -> distributions.py:33-41,83-92
  ValueNorm                   utils/valuenorm.py:32-79
  GAE returns                 buffer/shared_buffer.py:199-208
  advantage normalisation     algos/mappo.py:189-198
  PPO update (all quirks)     algos/mappo.py:103-187, utils/util.py:36-39 (one-sided Huber)
  grad clip + Adam            torch.nn.utils.clip_grad_norm_ / torch.optim.Adam as called at algos/mappo.py:176-185

Pinned by tests/test_oracle_mappo.py against tests/golden/mappo_*.npz, which tests/golden/make_golden_mappo.py
produced by running the UNMODIFIED reference learner (float32 torch) in the build container: forward values /
log-probs 1e-5, GAE returns 1e-5 relative, train_info and post-update parameters 2e-5.
"""
import numpy as np

LOG_2PI = float(np.log(2.0 * np.pi))
LN_EPS = 1e-5


# ---- layers -------------------------------------------------------------------------------------------
def ln_fwd(x, g, b):
    m = x.mean(-1, keepdims=True)
    v = ((x - m) ** 2).mean(-1, keepdims=True)
    r = 1.0 / np.sqrt(v + LN_EPS)
    xh = (x - m) * r
    return xh * g + b, (xh, r)


def ln_bwd(dy, cache, g):
    xh, r = cache
    dg = (dy * xh).sum(0)
    db = dy.sum(0)
    dxh = dy * g
    dx = r * (dxh - dxh.mean(-1, keepdims=True) - xh * (dxh * xh).mean(-1, keepdims=True))
    return dx, dg, db


class MLPNet:
    """feature_norm -> [Linear, ReLU, LayerNorm] x 2 -> head   (mlp.py:19-22,46-54; fc_h is never used)."""
    TRUNK = ("base.feature_norm.weight", "base.feature_norm.bias", "base.mlp.fc1.0.weight", "base.mlp.fc1.0.bias",
             "base.mlp.fc1.2.weight", "base.mlp.fc1.2.bias", "base.mlp.fc2.0.0.weight", "base.mlp.fc2.0.0.bias",
             "base.mlp.fc2.0.2.weight", "base.mlp.fc2.0.2.bias")

    def __init__(self, params, head_w, head_b):
        self.p = {k: np.asarray(v, dtype=np.float64).copy() for k, v in params.items()}
        self.head_w, self.head_b = head_w, head_b

    def names(self):
        return [k for k in self.p]

    def forward(self, x):
        p = self.p
        x = np.asarray(x, dtype=np.float64)
        h0, c0 = ln_fwd(x, p["base.feature_norm.weight"], p["base.feature_norm.bias"])
        z1 = h0 @ p["base.mlp.fc1.0.weight"].T + p["base.mlp.fc1.0.bias"]
        a1 = np.maximum(z1, 0)
        h1, c1 = ln_fwd(a1, p["base.mlp.fc1.2.weight"], p["base.mlp.fc1.2.bias"])
        z2 = h1 @ p["base.mlp.fc2.0.0.weight"].T + p["base.mlp.fc2.0.0.bias"]
        a2 = np.maximum(z2, 0)
        h2, c2 = ln_fwd(a2, p["base.mlp.fc2.0.2.weight"], p["base.mlp.fc2.0.2.bias"])
        out = h2 @ p[self.head_w].T + p[self.head_b]
        self.cache = (h0, c0, z1, h1, c1, z2, h2, c2)
        return out

    def backward(self, dout):
        p = self.p
        h0, c0, z1, h1, c1, z2, h2, c2 = self.cache
        g = {}
        g[self.head_w] = dout.T @ h2
        g[self.head_b] = dout.sum(0)
        dh2 = dout @ p[self.head_w]
        da2, g["base.mlp.fc2.0.2.weight"], g["base.mlp.fc2.0.2.bias"] = ln_bwd(dh2, c2, p["base.mlp.fc2.0.2.weight"])
        dz2 = da2 * (z2 > 0)
        g["base.mlp.fc2.0.0.weight"] = dz2.T @ h1
        g["base.mlp.fc2.0.0.bias"] = dz2.sum(0)
        dh1 = dz2 @ p["base.mlp.fc2.0.0.weight"]
        da1, g["base.mlp.fc1.2.weight"], g["base.mlp.fc1.2.bias"] = ln_bwd(dh1, c1, p["base.mlp.fc1.2.weight"])
        dz1 = da1 * (z1 > 0)
        g["base.mlp.fc1.0.weight"] = dz1.T @ h0
        g["base.mlp.fc1.0.bias"] = dz1.sum(0)
        dh0 = dz1 @ p["base.mlp.fc1.0.weight"]
        _, g["base.feature_norm.weight"], g["base.feature_norm.bias"] = ln_bwd(dh0, c0, p["base.feature_norm.weight"])
        return g


def make_actor(params):
    return MLPNet(params, "act.action_out.fc_mean.weight", "act.action_out.fc_mean.bias")


def make_critic(params):
    return MLPNet(params, "v_out.weight", "v_out.bias")


def gaussian_logp_entropy(mean, logstd, action):
    """FixedNormal.log_probs (sum over action dims) and the summed per-row entropy (act.py:172-184)."""
    std = np.exp(logstd)
    logp = (-((action - mean) ** 2) / (2 * std ** 2) - logstd - 0.5 * LOG_2PI).sum(-1, keepdims=True)
    ent = (0.5 + 0.5 * LOG_2PI + logstd).sum()
    return logp, ent


# ---- ValueNorm (utils/valuenorm.py) -----------------------------------------------------------------------
class ValueNorm:
    def __init__(self, state=(0.0, 0.0, 0.0), beta=0.99999, eps=1e-5):
        self.m, self.s, self.c = (float(x) for x in state)
        self.beta, self.eps = beta, eps

    def mean_var(self):
        c = max(self.c, self.eps)
        mean = self.m / c
        var = max(self.s / c - mean ** 2, 1e-2)
        return mean, var

    def update(self, x):
        x = np.asarray(x, dtype=np.float64)
        w = self.beta
        self.m = self.m * w + x.mean() * (1 - w)
        self.s = self.s * w + (x ** 2).mean() * (1 - w)
        self.c = self.c * w + (1 - w)

    def normalize(self, x):
        mean, var = self.mean_var()
        return (np.asarray(x, dtype=np.float64) - mean) / np.sqrt(var)

    def denormalize(self, x):
        mean, var = self.mean_var()
        return np.asarray(x, dtype=np.float64) * np.sqrt(var) + mean

    def state(self):
        return np.array([self.m, self.s, self.c])


# ---- GAE (buffer/shared_buffer.py:199-208) ----------------------------------------------------------------
def gae_returns(rewards, value_preds, masks, vn, gamma=0.99, lam=0.95):
    """rewards (T,...), value_preds (T+1,...) with value_preds[T] = next value (normalised units), masks (T+1,...)."""
    T = rewards.shape[0]
    den = vn.denormalize(value_preds)
    ret = np.zeros_like(den)
    gae = 0.0
    for t in reversed(range(T)):
        delta = rewards[t] + gamma * den[t + 1] * masks[t + 1] - den[t]
        gae = delta + gamma * lam * masks[t + 1] * gae
        ret[t] = gae + den[t]
    return ret


def huber(e, d):
    a = (np.abs(e) <= d).astype(np.float64)
    b = (e > d).astype(np.float64)          # one-sided on purpose (utils/util.py:36-39)
    return a * e ** 2 / 2 + b * d * (np.abs(e) - d / 2), a * e + b * d


class Adam:
    """torch.optim.Adam(lr, betas=(0.9, 0.999), eps, weight_decay=0) over a dict of float64 arrays."""

    def __init__(self, params, eps=1e-5, b1=0.9, b2=0.999):
        self.p, self.eps, self.b1, self.b2 = params, eps, b1, b2
        self.m = {k: np.zeros_like(v) for k, v in params.items()}
        self.v = {k: np.zeros_like(v) for k, v in params.items()}
        self.t = 0

    def step(self, grads, lr):
        self.t += 1
        bc1, bc2 = 1 - self.b1 ** self.t, 1 - self.b2 ** self.t
        for k, g in grads.items():
            self.m[k] = self.b1 * self.m[k] + (1 - self.b1) * g
            self.v[k] = self.b2 * self.v[k] + (1 - self.b2) * g * g
            self.p[k] -= (lr / bc1) * self.m[k] / (np.sqrt(self.v[k]) / np.sqrt(bc2) + self.eps)


def clip_grads(grads, max_norm):
    total = float(np.sqrt(sum((g.astype(np.float64) ** 2).sum() for g in grads.values())))
    coef = min(max_norm / (total + 1e-6), 1.0)
    for k in grads:
        grads[k] = grads[k] * coef
    return total


class Trainer:
    """MAPPOTrainer.train on a recorded rollout buffer (algos/mappo.py:189-227), one minibatch = all rows."""

    def __init__(self, actor_params, critic_params, hp, vn_state=(0.0, 0.0, 0.0)):
        self.actor, self.critic = make_actor(actor_params), make_critic(critic_params)
        self.hp = hp
        self.vn = ValueNorm(vn_state)
        self.opt_a = Adam(self.actor.p, eps=hp["opti_eps"])
        self.opt_c = Adam(self.critic.p, eps=hp["opti_eps"])

    def train(self, obs, actions, logp_old, value_preds, returns, lr, ppo_epoch):
        """obs (T+1,E,N,D); actions (T,E,N,2); logp_old (T,E,N,1); value_preds/returns (T+1,E,N,1)."""
        hp = self.hp
        T, E, N = actions.shape[:3]
        B = T * E * N
        c = hp["clip_param"]
        adv = returns[:-1].astype(np.float64) - self.vn.denormalize(value_preds[:-1])
        adv = (adv - adv.mean()) / (adv.std() + 1e-5)
        x = obs[:-1].reshape(B, -1).astype(np.float64)
        sx = np.repeat(obs[:-1].reshape(T * E, 1, -1), N, axis=1).reshape(B, -1).astype(np.float64)
        act = actions.reshape(B, -1).astype(np.float64)
        lpo = logp_old.reshape(B, 1).astype(np.float64)
        vold = value_preds[:-1].reshape(B, 1).astype(np.float64)
        ret = returns[:-1].reshape(B, 1).astype(np.float64)
        A = adv.reshape(B, 1)
        info = dict(value_loss=0.0, policy_loss=0.0, dist_entropy=0.0, actor_grad_norm=0.0, critic_grad_norm=0.0,
                    ratio=0.0)
        for _ in range(ppo_epoch):
            mean = self.actor.forward(x)
            logstd = self.actor.p["act.action_out.logstd._bias"].reshape(1, -1)
            logp, ent = gaussian_logp_entropy(mean, logstd, act)
            v = self.critic.forward(sx)
            ratio = np.exp(logp - lpo)
            s1, s2 = ratio * A, np.clip(ratio, 1 - c, 1 + c) * A
            policy_loss = -(2.0 * np.minimum(s1, s2)).mean()    # 2 equal log-prob columns (shared_buffer.py:61-62)
            inrange = (ratio >= 1 - c) & (ratio <= 1 + c)
            dmin = np.where(inrange, A, np.where(s1 < s2, A, 0.0))
            dlogp = -(2.0 / B) * dmin * ratio
            # value loss (cal_value_loss, mappo.py:103-131): ValueNorm.update first
            self.vn.update(ret)
            nret = self.vn.normalize(ret)
            e, dvc = nret - v, np.clip(v - vold, -c, c)
            e_c = nret - (vold + dvc)
            h, dh = huber(e, hp["huber_delta"])
            h_c, dh_c = huber(e_c, hp["huber_delta"])
            value_loss = np.maximum(h, h_c).mean()
            w1 = np.where(h > h_c, 1.0, np.where(h < h_c, 0.0, 0.5))
            in_v = (np.abs(v - vold) <= c).astype(np.float64)
            dv = (w1 * (-dh) + (1 - w1) * (-dh_c) * in_v) / B * hp["value_loss_coef"]
            # actor backward
            std2 = np.exp(2 * logstd)
            dmean = dlogp * (act - mean) / std2
            ga = self.actor.backward(dmean)
            ga["act.action_out.logstd._bias"] = ((dlogp * ((act - mean) ** 2 / std2 - 1.0)).sum(0)
                                                 - hp["entropy_coef"]).reshape(-1, 1)
            gc = self.critic.backward(dv)
            an = clip_grads(ga, hp["max_grad_norm"])
            cn = clip_grads(gc, hp["max_grad_norm"])
            self.opt_a.step(ga, lr)
            self.opt_c.step(gc, lr)
            info["value_loss"] += value_loss; info["policy_loss"] += policy_loss; info["dist_entropy"] += ent
            info["actor_grad_norm"] += an; info["critic_grad_norm"] += cn; info["ratio"] += ratio.mean()
        return {k: v / ppo_epoch for k, v in info.items()}
