"""NumPy restatement of the reference's MAPPO math.  TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/ (and bench.py's CPU-baseline leg) import this module; the product package never does.
It restates, in float64 NumPy with hand-written backward passes (paths relative to
/root/reference/uav_dcc_control/):

  actor / critic forward      algos/r_actor_critic.py:43-57,111-121 -> algos/algo_utils/mlp.py:25-29,52-58
                              -> algos/algo_utils/act.py:79-84,165-184 -> distributions.py:33-41,83-92
  ValueNorm                   utils/valuenorm.py:32-79
  GAE / discounted returns    buffer/shared_buffer.py:154-212 (use_gae, with / without a value normaliser)
  advantage normalisation     algos/mappo.py:189-198
  PPO update (all quirks)     algos/mappo.py:103-187, utils/util.py:36-43 (one-sided Huber / mse), with the
                              mappo.yaml switches use_huber_loss, use_clipped_value_loss, use_max_grad_norm,
                              use_valuenorm, weight_decay
  minibatches                 buffer/shared_buffer.py:219-279 (num_mini_batch index lists cut from a permutation)
  grad clip + Adam            torch.nn.utils.clip_grad_norm_ / torch.optim.Adam as called at algos/mappo.py:30-37,176-185
  recurrent policies          algos/algo_utils/rnn.py:8-80 (torch.nn.GRU, recurrent_N layers, + LayerNorm; hidden state times
                              the mask before every step), algos/r_actor_critic.py:55-57,118-120, and the two sequence
                              generators buffer/shared_buffer.py:281-376 (whole episodes) / :378-470 (data_chunk_length chunks)

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
    """feature_norm -> [Linear, act, LayerNorm] x (1 + layer_N) -> head   (mlp.py:13-29,44-58; fc_h is never used).
    act = "relu" | "tanh" (use_ReLU); the input LayerNorm is skipped when the parameters carry no
    base.feature_norm.* entries (use_feature_normalization: false)."""
    TRUNK = ("base.feature_norm.weight", "base.feature_norm.bias", "base.mlp.fc1.0.weight", "base.mlp.fc1.0.bias",
             "base.mlp.fc1.2.weight", "base.mlp.fc1.2.bias", "base.mlp.fc2.0.0.weight", "base.mlp.fc2.0.0.bias",
             "base.mlp.fc2.0.2.weight", "base.mlp.fc2.0.2.bias")

    def __init__(self, params, head_w, head_b, act="relu"):
        self.p = {k: np.asarray(v, dtype=np.float64).copy() for k, v in params.items()}
        self.head_w, self.head_b = head_w, head_b
        self.act = act
        self.fnorm = "base.feature_norm.weight" in self.p
        self.rnn_layers = 0          # recurrent_N GRU layers between the trunk and the head (rnn.py:8-22)
        while "rnn.rnn.weight_ih_l%d" % self.rnn_layers in self.p:
            self.rnn_layers += 1

    def _act(self, z):
        return np.maximum(z, 0) if self.act == "relu" else np.tanh(z)

    def _dact(self, da, z, a):
        return da * (z > 0) if self.act == "relu" else da * (1.0 - a * a)

    def names(self):
        return [k for k in self.p]

    def _blocks(self):
        """(Linear weight, bias, LayerNorm gain, bias) key stems of the hidden blocks: fc1, fc2.0, fc2.1, ... (layer_N)."""
        stems = ["base.mlp.fc1"]
        i = 0
        while "base.mlp.fc2.%d.0.weight" % i in self.p:
            stems.append("base.mlp.fc2.%d" % i)
            i += 1
        return stems

    def forward(self, x):
        p = self.p
        x = np.asarray(x, dtype=np.float64)
        if self.fnorm:
            h, c0 = ln_fwd(x, p["base.feature_norm.weight"], p["base.feature_norm.bias"])
        else:
            h, c0 = x, None
        self.c0, self.blocks = c0, []
        for st in self._blocks():
            z = h @ p[st + ".0.weight"].T + p[st + ".0.bias"]
            a = self._act(z)
            hn, c = ln_fwd(a, p[st + ".2.weight"], p[st + ".2.bias"])
            self.blocks.append((st, h, z, a, c))
            h = hn
        self.h_last = h
        if self.rnn_layers:          # recurrent nets: the caller runs the GRU on h_last, then head()
            return h
        return self.head(h)

    def head(self, y):
        self.y_head = y
        return y @ self.p[self.head_w].T + self.p[self.head_b]

    def head_backward(self, dout, g):
        g[self.head_w] = dout.T @ self.y_head
        g[self.head_b] = dout.sum(0)
        return dout @ self.p[self.head_w]

    def backward(self, dout, dh_last=None):
        """dout: gradient of the head output; recurrent nets pass dh_last (gradient of the trunk output) instead."""
        p = self.p
        g = {}
        dh = self.head_backward(dout, g) if dh_last is None else dh_last
        for st, h_in, z, a, c in reversed(self.blocks):
            da, g[st + ".2.weight"], g[st + ".2.bias"] = ln_bwd(dh, c, p[st + ".2.weight"])
            dz = self._dact(da, z, a)
            g[st + ".0.weight"] = dz.T @ h_in
            g[st + ".0.bias"] = dz.sum(0)
            dh = dz @ p[st + ".0.weight"]
        if self.fnorm:
            _, g["base.feature_norm.weight"], g["base.feature_norm.bias"] = ln_bwd(dh, self.c0, p["base.feature_norm.weight"])
        return g


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def gru_cell(p, l, x, h):
    """torch.nn.GRU cell of layer l: gate order (r, z, n) along the 3H rows of weight_ih / weight_hh."""
    Wi, Wh = p["rnn.rnn.weight_ih_l%d" % l], p["rnn.rnn.weight_hh_l%d" % l]
    bi, bh = p["rnn.rnn.bias_ih_l%d" % l], p["rnn.rnn.bias_hh_l%d" % l]
    H = Wh.shape[1]
    gi, gh = x @ Wi.T + bi, h @ Wh.T + bh
    r = _sigmoid(gi[:, :H] + gh[:, :H])
    z = _sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
    n = np.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
    hn = (1.0 - z) * n + z * h
    return hn, (x, h, r, z, n, gh[:, 2 * H:])


def gru_cell_bwd(p, l, dh, cache, g):
    """dh: gradient of the cell's new state.  Accumulates the layer's parameter gradients into g; returns (dx, dh_prev)."""
    x, h, r, z, n, ghn = cache
    Wi, Wh = p["rnn.rnn.weight_ih_l%d" % l], p["rnn.rnn.weight_hh_l%d" % l]
    dn = dh * (1.0 - z)
    dz = dh * (h - n)
    dnp = dn * (1.0 - n * n)
    dr = dnp * ghn
    drp, dzp = dr * r * (1.0 - r), dz * z * (1.0 - z)
    dgi = np.concatenate([drp, dzp, dnp], axis=1)
    dgh = np.concatenate([drp, dzp, dnp * r], axis=1)
    for k, v in (("rnn.rnn.weight_ih_l%d" % l, dgi.T @ x), ("rnn.rnn.weight_hh_l%d" % l, dgh.T @ h),
                 ("rnn.rnn.bias_ih_l%d" % l, dgi.sum(0)), ("rnn.rnn.bias_hh_l%d" % l, dgh.sum(0))):
        g[k] = g.get(k, 0.0) + v
    return dgi @ Wi, dh * z + dgh @ Wh


class RecurrentNet:
    """MLPBase -> RNNLayer (GRU x recurrent_N + LayerNorm) -> head  (r_actor_critic.py:43-57,111-121; rnn.py:24-80).
    Sequences are time-major: x (L*S, in) with row t*S + s, h0 (S, recurrent_N, H), masks (L*S, 1).  The hidden state is
    multiplied by the step's mask before every step — what rnn.py:38-69 does segment-wise (a mask of 1 is a no-op)."""

    def __init__(self, net):
        self.net, self.p = net, net.p

    def forward(self, x, h0, masks, L):
        net, p = self.net, self.p
        feats = net.forward(x)                       # trunk only (net.rnn_layers > 0)
        S = feats.shape[0] // L
        R = net.rnn_layers
        h = [np.asarray(h0, dtype=np.float64)[:, l].copy() for l in range(R)]
        masks = np.asarray(masks, dtype=np.float64).reshape(L, S, 1)
        self.caches, outs = [], []
        for t in range(L):
            inp = feats[t * S:(t + 1) * S]
            step = []
            for l in range(R):
                h[l], c = gru_cell(p, l, inp, h[l] * masks[t])
                step.append(c)
                inp = h[l]
            self.caches.append(step)
            outs.append(inp)
        self.masks, self.L, self.S = masks, L, S
        y, self.c_norm = ln_fwd(np.concatenate(outs, 0), p["rnn.norm.weight"], p["rnn.norm.bias"])
        self.h_final = np.stack(h, axis=1)           # (S, recurrent_N, H)
        return net.head(y)

    def backward(self, dout):
        net, p = self.net, self.p
        g = {}
        dy = net.head_backward(dout, g)
        dseq, g["rnn.norm.weight"], g["rnn.norm.bias"] = ln_bwd(dy, self.c_norm, p["rnn.norm.weight"])
        L, S, R = self.L, self.S, net.rnn_layers
        dh_next = [0.0] * R
        dfeat = np.zeros_like(net.h_last)
        for t in reversed(range(L)):
            dout_l = dseq[t * S:(t + 1) * S]             # gradient arriving at the top layer's output of step t
            for l in reversed(range(R)):
                dx, dhp = gru_cell_bwd(p, l, dout_l + dh_next[l], self.caches[t][l], g)
                dh_next[l] = dhp * self.masks[t]          # h_prev = h_{t-1} * mask_t
                dout_l = dx
            dfeat[t * S:(t + 1) * S] = dout_l
        g.update(net.backward(None, dh_last=dfeat))
        return g


def make_actor(params, act="relu"):
    return MLPNet(params, "act.action_out.fc_mean.weight", "act.action_out.fc_mean.bias", act)


def make_critic(params, act="relu"):
    return MLPNet(params, "v_out.weight", "v_out.bias", act)


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
def gae_returns(rewards, value_preds, masks, vn, gamma=0.99, lam=0.95, use_gae=True):
    """rewards (T,...), value_preds (T+1,...) with value_preds[T] = next value (normalised units), masks (T+1,...).
    vn None = no value normaliser.  use_gae False = shared_buffer.py:209-212: discounted returns bootstrapped from
    the RAW next value (returns[T] = next_value, never denormalised)."""
    T = rewards.shape[0]
    vp = np.asarray(value_preds, dtype=np.float64)
    if not use_gae:
        ret = np.zeros_like(vp)
        ret[T] = vp[T]
        for t in reversed(range(T)):
            ret[t] = ret[t + 1] * gamma * masks[t + 1] + rewards[t]
        return ret
    den = vn.denormalize(vp) if vn is not None else vp
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


def mse(e):
    return e ** 2 / 2, e


class Adam:
    """torch.optim.Adam(lr, betas=(0.9, 0.999), eps, weight_decay) over a dict of float64 arrays."""

    def __init__(self, params, eps=1e-5, b1=0.9, b2=0.999, weight_decay=0.0):
        self.p, self.eps, self.b1, self.b2, self.wd = params, eps, b1, b2, weight_decay
        self.m = {k: np.zeros_like(v) for k, v in params.items()}
        self.v = {k: np.zeros_like(v) for k, v in params.items()}
        self.t = 0

    def step(self, grads, lr):
        self.t += 1
        bc1, bc2 = 1 - self.b1 ** self.t, 1 - self.b2 ** self.t
        for k, g in grads.items():
            if self.wd:
                g = g + self.wd * self.p[k]
            self.m[k] = self.b1 * self.m[k] + (1 - self.b1) * g
            self.v[k] = self.b2 * self.v[k] + (1 - self.b2) * g * g
            self.p[k] -= (lr / bc1) * self.m[k] / (np.sqrt(self.v[k]) / np.sqrt(bc2) + self.eps)


def clip_grads(grads, max_norm, do_clip=True):
    total = float(np.sqrt(sum((g.astype(np.float64) ** 2).sum() for g in grads.values())))
    coef = min(max_norm / (total + 1e-6), 1.0) if do_clip else 1.0   # get_gard_norm only reports (mappo.py:179-181)
    for k in grads:
        grads[k] = grads[k] * coef
    return total


class Trainer:
    """MAPPOTrainer.train on a recorded rollout buffer (algos/mappo.py:189-227).  hp may carry the mappo.yaml
    switches use_huber_loss / use_clipped_value_loss / use_max_grad_norm / use_valuenorm / use_ReLU /
    use_centralized_V (default True), weight_decay (0) and num_mini_batch (1)."""

    def __init__(self, actor_params, critic_params, hp, vn_state=(0.0, 0.0, 0.0)):
        act = "relu" if hp.get("use_ReLU", True) else "tanh"
        self.actor, self.critic = make_actor(actor_params, act), make_critic(critic_params, act)
        self.hp = hp
        self.vn = ValueNorm(vn_state) if hp.get("use_valuenorm", True) else None
        wd = float(hp.get("weight_decay", 0.0))
        self.opt_a = Adam(self.actor.p, eps=hp["opti_eps"], weight_decay=wd)
        self.opt_c = Adam(self.critic.p, eps=hp["opti_eps"], weight_decay=wd)

    def train(self, obs, actions, logp_old, value_preds, returns, lr, ppo_epoch, perms=None, rnn_states=None,
              rnn_states_critic=None, masks=None):
        """obs (T+1,E,N,D); actions (T,E,N,2); logp_old (T,E,N,1); value_preds/returns (T+1,E,N,1).
        perms (num_mini_batch > 1): per epoch, the permutation of the T*E*N agent rows the generator drew.
        Recurrent policies (use_recurrent_policy / use_naive_recurrent_policy): rnn_states / rnn_states_critic
        (T+1,E,N,recurrent_N,H) and masks (T+1,E,N,1) as the rollout stored them; perms = per epoch the permutation of the
        sequence chunks (shared_buffer.py:391 / :294).  Both generators cut chunks of L consecutive entries out of the
        rollout flattened in (env, agent, time) order — L = data_chunk_length (chunked) or T (naive: whole episodes) —
        start each chunk from the stored hidden state of its first entry and feed them time-major (L, chunks)."""
        hp = self.hp
        T, E, N = actions.shape[:3]
        B = T * E * N
        recurrent = bool(hp.get("use_recurrent_policy", False) or hp.get("use_naive_recurrent_policy", False))
        if recurrent:
            L = int(hp["data_chunk_length"]) if hp.get("use_recurrent_policy", False) else T
            n_chunks = B // L
            ract, rcrit = RecurrentNet(self.actor), RecurrentNet(self.critic)
            R, Hh = rnn_states.shape[3:]
            hs_a = np.asarray(rnn_states[:-1], dtype=np.float64).reshape(B, R, Hh)
            hs_c = np.asarray(rnn_states_critic[:-1], dtype=np.float64).reshape(B, R, Hh)
            mk_all = np.asarray(masks[:-1], dtype=np.float64).reshape(B, 1)
        c = hp["clip_param"]
        nmb = int(hp.get("num_mini_batch", 1))
        vp = value_preds[:-1].astype(np.float64)
        adv = returns[:-1].astype(np.float64) - (self.vn.denormalize(vp) if self.vn is not None else vp)
        adv = (adv - adv.mean()) / (adv.std() + 1e-5)
        x_all = obs[:-1].reshape(B, -1).astype(np.float64)
        if hp.get("use_centralized_V", True):
            sx_all = np.repeat(obs[:-1].reshape(T * E, 1, -1), N, axis=1).reshape(B, -1).astype(np.float64)
        else:                       # decentralised critic: share_obs = obs (learner.py:221-222,272-273)
            sx_all = x_all
        act_all = actions.reshape(B, -1).astype(np.float64)
        lpo_all = logp_old.reshape(B, 1).astype(np.float64)
        vold_all = value_preds[:-1].reshape(B, 1).astype(np.float64)
        ret_all = returns[:-1].reshape(B, 1).astype(np.float64)
        A_all = adv.reshape(B, 1)
        loss_fn = (lambda e: huber(e, hp["huber_delta"])) if hp.get("use_huber_loss", True) else mse
        clipped = hp.get("use_clipped_value_loss", True)
        info = dict(value_loss=0.0, policy_loss=0.0, dist_entropy=0.0, actor_grad_norm=0.0, critic_grad_norm=0.0,
                    ratio=0.0)
        mbs = (n_chunks if recurrent else B) // nmb
        for ep in range(ppo_epoch):
            for i in range(nmb):
                if recurrent:
                    ch = np.asarray(perms[ep][i * mbs:(i + 1) * mbs], dtype=np.int64)
                    f = ch[None, :] * L + np.arange(L)[:, None]            # (L, chunks): entries in (env, agent, time) order
                    ea, t = f // T, f % T
                    rows = t * (E * N) + (ea // N) * N + ea % N            # the same entries as rows of the (T, E, N) arrays
                    sel = rows.reshape(-1)
                    Bm = sel.size
                elif nmb == 1:
                    sel = slice(None)      # a permutation of the whole batch only reorders the sums
                    Bm = B
                else:
                    sel = np.asarray(perms[ep][i * mbs:(i + 1) * mbs], dtype=np.int64)
                    Bm = mbs
                x, sx, act, lpo, vold, ret, A = (a[sel] for a in (x_all, sx_all, act_all, lpo_all, vold_all, ret_all, A_all))
                if recurrent:
                    mean = ract.forward(x, hs_a[rows[0]], mk_all[sel], L)
                else:
                    mean = self.actor.forward(x)
                logstd = self.actor.p["act.action_out.logstd._bias"].reshape(1, -1)
                logp, ent = gaussian_logp_entropy(mean, logstd, act)
                v = rcrit.forward(sx, hs_c[rows[0]], mk_all[sel], L) if recurrent else self.critic.forward(sx)
                ratio = np.exp(logp - lpo)
                s1, s2 = ratio * A, np.clip(ratio, 1 - c, 1 + c) * A
                policy_loss = -(2.0 * np.minimum(s1, s2)).mean()    # 2 equal log-prob columns (shared_buffer.py:61-62)
                inrange = (ratio >= 1 - c) & (ratio <= 1 + c)
                dmin = np.where(inrange, A, np.where(s1 < s2, A, 0.0))
                dlogp = -(2.0 / Bm) * dmin * ratio
                # value loss (cal_value_loss, mappo.py:103-131): ValueNorm.update first
                if self.vn is not None:
                    self.vn.update(ret)
                    nret = self.vn.normalize(ret)
                else:
                    nret = ret
                e, dvc = nret - v, np.clip(v - vold, -c, c)
                e_c = nret - (vold + dvc)
                h, dh = loss_fn(e)
                h_c, dh_c = loss_fn(e_c)
                if clipped:
                    value_loss = np.maximum(h, h_c).mean()
                    w1 = np.where(h > h_c, 1.0, np.where(h < h_c, 0.0, 0.5))
                else:
                    value_loss = h.mean()
                    w1 = np.ones_like(h)
                in_v = (np.abs(v - vold) <= c).astype(np.float64)
                dv = (w1 * (-dh) + (1 - w1) * (-dh_c) * in_v) / Bm * hp["value_loss_coef"]
                # actor backward
                std2 = np.exp(2 * logstd)
                dmean = dlogp * (act - mean) / std2
                ga = ract.backward(dmean) if recurrent else self.actor.backward(dmean)
                ga["act.action_out.logstd._bias"] = ((dlogp * ((act - mean) ** 2 / std2 - 1.0)).sum(0)
                                                     - hp["entropy_coef"]).reshape(-1, 1)
                gc = rcrit.backward(dv) if recurrent else self.critic.backward(dv)
                do_clip = hp.get("use_max_grad_norm", True)
                an = clip_grads(ga, hp["max_grad_norm"], do_clip)
                cn = clip_grads(gc, hp["max_grad_norm"], do_clip)
                self.opt_a.step(ga, lr)
                self.opt_c.step(gc, lr)
                info["value_loss"] += value_loss; info["policy_loss"] += policy_loss; info["dist_entropy"] += ent
                info["actor_grad_norm"] += an; info["critic_grad_norm"] += cn; info["ratio"] += ratio.mean()
        return {k: v / (ppo_epoch * nmb) for k, v in info.items()}
