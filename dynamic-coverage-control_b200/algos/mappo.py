"""MAPPOPolicy / MAPPOTrainer — the reference's algo boundary (algos/mappo.py:15-247) on the CUDA learner kernels.

Same class names, method names, argument order and return values as the reference, so `Learner` reads the same;
the bodies call the C ABI (`dcc_mappo_*`, include/dcc_b200.h) instead of torch.nn / autograd:

  MAPPOPolicy.get_actions / get_values / act / evaluate_actions  -> dcc_mappo_act / dcc_mappo_evaluate
  MAPPOPolicy.lr_decay                                           -> utils/util.py:29-33 (a float, no kernel)
  MAPPOTrainer.train                                             -> dcc_mappo_train_begin, then per epoch
                                                                    dcc_mappo_epoch_grads [+ all-reduce] + dcc_mappo_apply x2
  recurrent policies (use_recurrent_policy / use_naive_recurrent_policy: GRU x recurrent_N + LayerNorm between trunk
  and head, algos/algo_utils/rnn.py)                             -> dcc_mappo_act_rnn, dcc_mappo_seq_grads

Parameters live in flat float32 CUDA tensors (layout in include/dcc_b200.h); `state_dict()` / `load_state_dict()`
speak the reference's key names so weights round-trip with reference checkpoints.
Tensors in, tensors out: no numpy on the hot path.  There is no fallback: a missing library raises DccError.
"""
import ctypes as C
import math
import os
import pickle
from collections import OrderedDict

import numpy as np
import torch

from .. import _lib
from ..parallel import Comm
from .valuenorm import ValueNorm

TRUNK = ("base.feature_norm.weight", "base.feature_norm.bias", "base.mlp.fc1.0.weight", "base.mlp.fc1.0.bias",
         "base.mlp.fc1.2.weight", "base.mlp.fc1.2.bias", "base.mlp.fc2.0.0.weight", "base.mlp.fc2.0.0.bias",
         "base.mlp.fc2.0.2.weight", "base.mlp.fc2.0.2.bias")      # the shipped layer_N = 1 trunk


def trunk_keys(layer_N=1):
    """state_dict keys of MLPBase for `layer_N` fc2 blocks (mlp.py:16-29), without the never-used fc_h block."""
    keys = list(TRUNK[:6])
    for i in range(layer_N):
        keys += ["base.mlp.fc2.%d.0.weight" % i, "base.mlp.fc2.%d.0.bias" % i, "base.mlp.fc2.%d.2.weight" % i,
                 "base.mlp.fc2.%d.2.bias" % i]
    return keys
FC_H = ("base.mlp.fc_h.0.weight", "base.mlp.fc_h.0.bias", "base.mlp.fc_h.2.weight", "base.mlp.fc_h.2.bias")


def rnn_keys(recurrent_N, hidden):
    """state_dict keys / shapes of the RNNLayer (rnn.py:13-22): torch.nn.GRU's per-layer tensors, then the LayerNorm."""
    out = []
    for i in range(recurrent_N):
        out += [("rnn.rnn.weight_ih_l%d" % i, (3 * hidden, hidden)), ("rnn.rnn.weight_hh_l%d" % i, (3 * hidden, hidden)),
                ("rnn.rnn.bias_ih_l%d" % i, (3 * hidden,)), ("rnn.rnn.bias_hh_l%d" % i, (3 * hidden,))]
    if recurrent_N:
        out += [("rnn.norm.weight", (hidden,)), ("rnn.norm.bias", (hidden,))]
    return out


def net_layout(in_dim, hidden, out_dim, head, logstd=False, feature_norm=True, layer_N=1, recurrent_N=0):
    """name -> (offset, shape) of the flat parameter buffer, in the order include/dcc_b200.h documents.
    feature_norm=False (use_feature_normalization: false): the net has no base.feature_norm.* entries (mlp.py:44-45).
    recurrent_N > 0: the RNNLayer's tensors sit between the trunk and the head (r_actor_critic.py:33-39)."""
    shapes = [(in_dim,), (in_dim,), (hidden, in_dim), (hidden,), (hidden,), (hidden,)] + \
        [(hidden, hidden), (hidden,), (hidden,), (hidden,)] * layer_N
    lay, off = OrderedDict(), 0
    for k, shp in zip(trunk_keys(layer_N), shapes):
        if not feature_norm and k.startswith("base.feature_norm"):
            continue
        lay[k] = (off, shp)
        off += int(np.prod(shp))
    for k, shp in rnn_keys(recurrent_N, hidden):
        lay[k] = (off, shp)
        off += int(np.prod(shp))
    lay[head + ".weight"] = (off, (out_dim, hidden)); off += out_dim * hidden
    lay[head + ".bias"] = (off, (out_dim,)); off += out_dim
    if logstd:
        lay["act.action_out.logstd._bias"] = (off, (out_dim, 1)); off += out_dim
    return lay, off


def _reference_init(in_dim, hidden, out_dim, head_gain, use_orthogonal=True, use_relu=True, feature_norm=True, layer_N=1,
                    recurrent_N=0):
    """Initial parameters drawn exactly as the reference constructs a net (same torch RNG consumption order):
    MLPBase -> LayerNorm, fc1 = Linear + orthogonal / xavier_uniform (`use_orthogonal`) with the gain of the trunk
    activation (sqrt 2 for ReLU, 5/3 for tanh), fc_h likewise, fc2 = deepcopy(fc_h) (algos/algo_utils/mlp.py:13-23),
    then the head Linear with the same init method and `head_gain` (distributions.py:76-80, r_actor_critic.py:92-107).
    Host-side torch, construction time only."""
    import torch.nn as nn
    gain = nn.init.calculate_gain("relu" if use_relu else "tanh")
    init_method = nn.init.orthogonal_ if use_orthogonal else nn.init.xavier_uniform_

    def lin(i, o, g):
        m = nn.Linear(i, o)
        init_method(m.weight.data, gain=g)
        nn.init.constant_(m.bias.data, 0)
        return m
    fc1 = lin(in_dim, hidden, gain)
    fc_h = lin(hidden, hidden, gain)
    rnn = None
    if recurrent_N:      # RNNLayer is built between the trunk and the head (r_actor_critic.py:33-39): nn.GRU draws its default
        rnn = nn.GRU(hidden, hidden, num_layers=recurrent_N)     # init, then biases -> 0, weights -> orthogonal / xavier (rnn.py:13-21)
        for name, param in rnn.named_parameters():
            if "bias" in name:
                nn.init.constant_(param, 0)
            elif "weight" in name:
                init_method(param)
    head = lin(hidden, out_dim, head_gain)
    ones, zeros = torch.ones, torch.zeros
    sd = OrderedDict()
    if feature_norm:
        sd["base.feature_norm.weight"], sd["base.feature_norm.bias"] = ones(in_dim), zeros(in_dim)
    sd["base.mlp.fc1.0.weight"], sd["base.mlp.fc1.0.bias"] = fc1.weight.data.clone(), fc1.bias.data.clone()
    sd["base.mlp.fc1.2.weight"], sd["base.mlp.fc1.2.bias"] = ones(hidden), zeros(hidden)
    sd["base.mlp.fc_h.0.weight"], sd["base.mlp.fc_h.0.bias"] = fc_h.weight.data.clone(), fc_h.bias.data.clone()
    sd["base.mlp.fc_h.2.weight"], sd["base.mlp.fc_h.2.bias"] = ones(hidden), zeros(hidden)
    for i in range(layer_N):     # get_clones(fc_h, layer_N): every fc2 block starts as a copy of fc_h (mlp.py:23)
        sd["base.mlp.fc2.%d.0.weight" % i], sd["base.mlp.fc2.%d.0.bias" % i] = fc_h.weight.data.clone(), fc_h.bias.data.clone()
        sd["base.mlp.fc2.%d.2.weight" % i], sd["base.mlp.fc2.%d.2.bias" % i] = ones(hidden), zeros(hidden)
    if rnn is not None:
        for name, param in rnn.named_parameters():
            sd["rnn.rnn." + name] = param.data.clone()
        sd["rnn.norm.weight"], sd["rnn.norm.bias"] = ones(hidden), zeros(hidden)
    return sd, head


class _Net:
    """Flat parameter / gradient / Adam-moment storage of one network + its named views."""

    def __init__(self, layout, total, grad_storage, device):
        self.layout, self.total = layout, total
        self.params = torch.zeros(total, dtype=torch.float32, device=device)
        self.grads = grad_storage
        self.adam_m = torch.zeros(total, dtype=torch.float32, device=device)
        self.adam_v = torch.zeros(total, dtype=torch.float32, device=device)
        self.adam_step = 0
        self.fc_h = {}   # the never-used fc_h block (mlp.py:21-23): kept only so checkpoints round-trip

    def view(self, name, which="params"):
        off, shp = self.layout[name]
        return getattr(self, which)[off:off + int(np.prod(shp))].view(*shp)

    def state_dict(self):
        sd = OrderedDict()
        for k in self.layout:
            sd[k] = self.view(k).detach().clone()
            if k == "base.mlp.fc1.2.bias":
                for kk in FC_H:
                    if kk in self.fc_h:
                        sd[kk] = self.fc_h[kk].clone()
        return sd

    def load_state_dict(self, sd, strict=True):
        for k in self.layout:
            if k not in sd:
                if strict:
                    raise KeyError("missing key %s" % k)
                continue
            v = torch.as_tensor(np.asarray(sd[k].detach().cpu() if isinstance(sd[k], torch.Tensor) else sd[k]),
                                dtype=torch.float32)
            off, shp = self.layout[k]
            if int(v.numel()) != int(np.prod(shp)):
                raise ValueError("shape mismatch for %s: %s vs %s" % (k, tuple(v.shape), shp))
            self.view(k).copy_(v.reshape(*shp))
        for kk in FC_H:
            if kk in sd:
                self.fc_h[kk] = torch.as_tensor(np.asarray(sd[kk].detach().cpu() if isinstance(sd[kk], torch.Tensor)
                                                           else sd[kk]), dtype=torch.float32)


class MAPPOPolicy:
    """algos/mappo.py:15-65.  obs_space / cent_obs_space / act_space: objects with `.shape` (gym-like Box)."""

    def __init__(self, cfg, obs_space, cent_obs_space, act_space, device=None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.DccError("MAPPOPolicy needs a CUDA device (sm_100); there is no CPU fallback")
        if act_space.__class__.__name__ != "Box":
            raise NotImplementedError("only the continuous Box(2) action space of the env is implemented")
        from ..utils.config import check_supported
        check_supported(cfg)
        self.device = torch.device("cuda", int(getattr(cfg, "device", 0) or 0)) if device is None else torch.device(device)
        self.actor_lr, self.critic_lr = float(cfg.actor_lr), float(cfg.critic_lr)
        self.opti_eps, self.weight_decay = float(cfg.opti_eps), float(cfg.weight_decay)
        self.obs_space, self.share_obs_space, self.act_space = obs_space, cent_obs_space, act_space
        self.obs_dim = int(obs_space.shape[0])
        self.share_dim = int(cent_obs_space.shape[0])
        if self.share_dim % self.obs_dim:
            raise ValueError("centralised obs dim %d is not a multiple of obs dim %d" % (self.share_dim, self.obs_dim))
        # use_centralized_V: false hands obs_space in as cent_obs_space (learner.py:43-46): the critic then sees one agent's
        # observation, i.e. the kernels run with "1 agent per env" over E*N pseudo-envs and values are per agent
        self.n_agents = self.share_dim // self.obs_dim
        self.hidden = int(cfg.algo_hidden_size)
        self.act_dim = int(act_space.shape[0])

        mc = _lib.MappoCfg()
        _lib.check(self.lib.dcc_mappo_cfg_default(C.byref(mc)), "dcc_mappo_cfg_default")
        mc.n_agents, mc.obs_dim, mc.hidden, mc.act_dim = self.n_agents, self.obs_dim, self.hidden, self.act_dim
        mc.chunk_rows = int(getattr(cfg, "chunk_rows", 0) or 0)
        mc.gemm_backend = int(getattr(cfg, "gemm_backend", 0) or 0)
        mc.clip_param, mc.entropy_coef = float(cfg.clip_param), float(cfg.entropy_coef)
        mc.value_loss_coef, mc.huber_delta = float(cfg.value_loss_coef), float(cfg.huber_delta)
        mc.max_grad_norm, mc.gamma, mc.gae_lambda = float(cfg.max_grad_norm), float(cfg.gamma), float(cfg.gae_lambda)
        mc.opti_eps = self.opti_eps
        mc.use_huber_loss = 1 if getattr(cfg, "use_huber_loss", True) else 0
        mc.use_clipped_value_loss = 1 if getattr(cfg, "use_clipped_value_loss", True) else 0
        mc.use_max_grad_norm = 1 if getattr(cfg, "use_max_grad_norm", True) else 0
        mc.use_valuenorm = 1 if getattr(cfg, "use_valuenorm", True) else 0
        mc.use_gae = 1 if getattr(cfg, "use_gae", True) else 0
        mc.weight_decay = self.weight_decay
        fnorm, relu = bool(getattr(cfg, "use_feature_normalization", True)), bool(getattr(cfg, "use_ReLU", True))
        mc.use_feature_normalization, mc.use_relu = int(fnorm), int(relu)
        layer_N = int(getattr(cfg, "layer_N", 1))
        mc.layer_N = layer_N
        from ..utils.config import is_recurrent
        self.recurrent_N = int(getattr(cfg, "recurrent_N", 1)) if is_recurrent(cfg) else 0
        mc.recurrent_N = self.recurrent_N
        self.mcfg = mc
        h = C.c_void_p()
        _lib.check(self.lib.dcc_mappo_create(C.byref(mc), self.device.index, C.byref(h)), "dcc_mappo_create")
        self._h = h

        la, na = net_layout(self.obs_dim, self.hidden, self.act_dim, "act.action_out.fc_mean", logstd=True, feature_norm=fnorm,
                            layer_N=layer_N, recurrent_N=self.recurrent_N)
        lc, nc = net_layout(self.share_dim, self.hidden, 1, "v_out", feature_norm=fnorm, layer_N=layer_N,
                            recurrent_N=self.recurrent_N)
        assert na == self.lib.dcc_mappo_param_count(h, 0) and nc == self.lib.dcc_mappo_param_count(h, 1)
        # one flat gradient buffer for both nets: a single all-reduce per PPO epoch
        self.flat_grads = torch.zeros(na + nc, dtype=torch.float32, device=self.device)
        self.actor = _Net(la, na, self.flat_grads[:na], self.device)
        self.critic = _Net(lc, nc, self.flat_grads[na:], self.device)
        # initial weights: the reference's construction order — actor first, then critic (mappo.py:27-28)
        init_kw = dict(use_orthogonal=bool(getattr(cfg, "use_orthogonal", True)), use_relu=relu, feature_norm=fnorm,
                       layer_N=layer_N, recurrent_N=self.recurrent_N)
        sd, head = _reference_init(self.obs_dim, self.hidden, self.act_dim, float(cfg.gain), **init_kw)
        sd["act.action_out.fc_mean.weight"], sd["act.action_out.fc_mean.bias"] = head.weight.data, head.bias.data
        sd["act.action_out.logstd._bias"] = torch.zeros(self.act_dim, 1)
        self.actor.load_state_dict(sd)
        sd, head = _reference_init(self.share_dim, self.hidden, 1, 1.0, **init_kw)
        sd["v_out.weight"], sd["v_out.bias"] = head.weight.data, head.bias.data
        self.critic.load_state_dict(sd)

        self.lr_actor_now, self.lr_critic_now = self.actor_lr, self.critic_lr
        self.seed = int(getattr(cfg, "seed", 0))
        self._rng_offset = 0
        self._scratch = {}

    # ---- helpers -----------------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _ptr(t):
        return None if t is None else C.c_void_p(t.data_ptr())

    def _out(self, key, shape):
        t = self._scratch.get(key)
        if t is None or tuple(t.shape) != tuple(shape):
            t = torch.empty(shape, dtype=torch.float32, device=self.device)
            self._scratch[key] = t
        return t

    def _as_obs(self, obs):
        if not (isinstance(obs, torch.Tensor) and obs.is_cuda and obs.dtype == torch.float32):
            obs = torch.as_tensor(np.asarray(obs), dtype=torch.float32).to(self.device)
        obs = obs.contiguous()
        n = obs.numel() // (self.n_agents * self.obs_dim)
        if n * self.n_agents * self.obs_dim != obs.numel():
            raise ValueError("obs has %d elements, not a multiple of N*D = %d" % (obs.numel(), self.n_agents * self.obs_dim))
        return obs, n

    def launch_count(self):
        return int(self.lib.dcc_mappo_launch_count(self._h))

    def gemm_backend(self):
        # backend 2: tcgen05 GEMMs in split precision at fp32-FFMA accuracy — fp16 hi/lo split (forward, dX and weight-gradient
        # GEMMs fed by TMA from pre-split operands) with 3xTF32 where an operand is not range-bounded (raw observations);
        # the DCC_TC_* knobs move individual GEMMs back to 3xTF32
        return {1: "simt-fp32", 2: "tcgen05-split-fp16/3xtf32"}[int(self.lib.dcc_mappo_gemm_backend(self._h))]

    # ---- reference surface ---------------------------------------------------------------------------------
    def lr_decay(self, episode, episodes):
        """update_linear_schedule (utils/util.py:29-33) for both optimisers."""
        self.lr_actor_now = max(self.actor_lr - self.actor_lr * (episode / float(episodes)), 0.0)
        self.lr_critic_now = max(self.critic_lr - self.critic_lr * (episode / float(episodes)), 0.0)

    def get_actions(self, cent_obs, obs, rnn_states_actor=None, rnn_states_critic=None, masks=None,
                    available_actions=None, deterministic=False, out_actions=None, out_logp=None, out_values=None,
                    out_rnn_actor=None, out_rnn_critic=None):
        """obs: (E*N, D) or (E, N, D) CUDA float32, the env's observation buffer.  cent_obs is accepted for signature
        parity and NOT read: the centralised input of env e is its N obs rows concatenated (learner.py:219-220),
        i.e. the same memory, evaluated once per env instead of N identical times.
        Returns (values (E*N,1), actions (E*N,2), action_log_probs (E*N,1), rnn_states_actor, rnn_states_critic)."""
        obs, n = self._as_obs(obs)
        N = self.n_agents
        actions = out_actions if out_actions is not None else self._out("actions", (n * N, 2))
        logp = out_logp if out_logp is not None else self._out("logp", (n * N,))
        values = out_values if out_values is not None else self._out("values", (n,))
        self._rng_offset += 1
        if self.recurrent_N:
            ha, hc = self._act_rnn(obs, n, rnn_states_actor, rnn_states_critic, masks, 0, deterministic, actions, logp, values,
                                   out_rnn_actor, out_rnn_critic)
            v = values.view(n, 1).expand(n, N).reshape(n * N, 1)
            return v, actions.view(n * N, 2), logp.view(n * N, 1), ha, hc
        _lib.check(self.lib.dcc_mappo_act(self._h, self._ptr(self.actor.params), self._ptr(self.critic.params),
                                          self._ptr(obs), n, self.seed, self._rng_offset, 1 if deterministic else 0,
                                          self._ptr(actions), self._ptr(logp), self._ptr(values), self._stream()),
                   "dcc_mappo_act")
        v = values.view(n, 1).expand(n, N).reshape(n * N, 1)
        return v, actions.view(n * N, 2), logp.view(n * N, 1), rnn_states_actor, rnn_states_critic

    # ---- recurrent policies -------------------------------------------------------------------------------
    def _per_env(self, x, n, tail):
        """(n*N, *tail) as the reference passes it (N identical rows per env), or (n, *tail) -> contiguous (n, *tail)."""
        x = torch.as_tensor(x, dtype=torch.float32, device=self.device) if not isinstance(x, torch.Tensor) else x.to(self.device, torch.float32)
        per = int(np.prod(tail)) if tail else 1
        if x.numel() == n * per:
            return x.reshape(n, *tail).contiguous()
        if x.numel() == n * self.n_agents * per:
            return x.reshape(n, self.n_agents, *tail)[:, 0].contiguous()
        raise ValueError("expected %d or %d rows of %s, got %d elements" % (n, n * self.n_agents, tail, x.numel()))

    def _act_rnn(self, obs, n, h_actor, h_critic, masks, mode, deterministic, actions, logp, values, out_ha=None, out_hc=None,
                 do_actor=True, do_critic=True):
        """One vec-env step of the recurrent nets (dcc_mappo_act_rnn).  h_actor (n*N, recurrent_N, H); h_critic
        (n*N, ...) as the reference stores it or (n, ...) one per env; masks (n*N, 1) or (n,).  Returns the new states:
        actor (n*N, recurrent_N, H) and critic expanded to the reference's (n*N, recurrent_N, H) (a view: one row per env)."""
        N, R, H = self.n_agents, self.recurrent_N, self.hidden
        if masks is None:
            raise ValueError("recurrent policies need the step's masks")
        mk = self._per_env(masks, n, ())
        ha = hc = None
        if do_actor:
            ha = torch.as_tensor(h_actor, dtype=torch.float32, device=self.device).reshape(n * N, R, H).contiguous()
            out_ha = out_ha if out_ha is not None else torch.empty((n * N, R, H), dtype=torch.float32, device=self.device)
        if do_critic:
            hc = self._per_env(h_critic, n, (R, H))
            out_hc = out_hc if out_hc is not None else torch.empty((n, R, H), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.dcc_mappo_act_rnn(
            self._h, self._ptr(self.actor.params) if do_actor else None, self._ptr(self.critic.params) if do_critic else None,
            self._ptr(obs), n, self._ptr(ha), self._ptr(hc), self._ptr(mk), int(mode), self.seed, self._rng_offset,
            1 if deterministic else 0, self._ptr(actions) if do_actor else None, self._ptr(logp) if do_actor else None,
            self._ptr(values) if do_critic else None, self._ptr(out_ha) if do_actor else None,
            self._ptr(out_hc) if do_critic else None, self._stream()), "dcc_mappo_act_rnn")
        new_ha = out_ha.view(n * N, R, H) if do_actor else None
        new_hc = out_hc.view(n, 1, R, H).expand(n, N, R, H).reshape(n * N, R, H) if do_critic else None
        return new_ha, new_hc

    def get_values(self, cent_obs, rnn_states_critic=None, masks=None, out_values=None, rows_repeated=False):
        """cent_obs: (E, N*D), or the obs buffer (E, N, D) — one centralised row per env.  rows_repeated=True takes the
        reference's layout instead, (E*N, N*D) with N identical rows per env (learner.py:281-285), and evaluates
        every N-th row.  Returns (E*N, 1), the reference's shape."""
        N, S = self.n_agents, self.share_dim
        if not (isinstance(cent_obs, torch.Tensor) and cent_obs.is_cuda and cent_obs.dtype == torch.float32):
            cent_obs = torch.as_tensor(np.asarray(cent_obs), dtype=torch.float32).to(self.device)
        if rows_repeated:
            cent_obs = cent_obs.reshape(-1, S)[::N]
        cent_obs = cent_obs.contiguous()
        n = cent_obs.numel() // S
        if n * S != cent_obs.numel():
            raise ValueError("cent_obs has %d elements, not a multiple of N*D = %d" % (cent_obs.numel(), S))
        values = out_values if out_values is not None else self._out("values", (n,))
        if self.recurrent_N:
            self._act_rnn(cent_obs, n, None, rnn_states_critic, masks, 0, False, None, None, values, do_actor=False)
            return values.view(n, 1).expand(n, N).reshape(n * N, 1)
        _lib.check(self.lib.dcc_mappo_act(self._h, None, self._ptr(self.critic.params), self._ptr(cent_obs), n, 0, 0, 0,
                                          None, None, self._ptr(values), self._stream()), "dcc_mappo_act(values)")
        return values.view(n, 1).expand(n, N).reshape(n * N, 1)

    def evaluate_actions(self, cent_obs, obs, rnn_states_actor, rnn_states_critic, action, masks=None,
                         available_actions=None, active_masks=None):
        """Forward-only evaluation: returns (values (B,1), action_log_probs (B,1), dist_entropy scalar tensor)."""
        obs, n = self._as_obs(obs)
        N = self.n_agents
        action = torch.as_tensor(action, dtype=torch.float32, device=self.device).contiguous()
        logp = self._out("ev_logp", (n * N,))
        values = self._out("ev_values", (n,))
        if self.recurrent_N:
            # one step per row (the x.size(0) == hxs.size(0) branch of RNNLayer.forward, rnn.py:25-29); sequences are
            # evaluated by the update itself (dcc_mappo_seq_grads)
            if rnn_states_actor is None or int(np.prod(tuple(rnn_states_actor.shape))) != n * N * self.recurrent_N * self.hidden:
                raise NotImplementedError("evaluate_actions on recurrent policies takes one hidden state per observation row")
            self._act_rnn(obs, n, rnn_states_actor, rnn_states_critic, masks, 1, False, action, logp, values)
        else:
            _lib.check(self.lib.dcc_mappo_evaluate(self._h, self._ptr(self.actor.params), self._ptr(self.critic.params),
                                                   self._ptr(obs), self._ptr(action), n, self._ptr(logp), self._ptr(values),
                                                   None, self._stream()), "dcc_mappo_evaluate")
        logstd = self.actor.view("act.action_out.logstd._bias")
        ent = (0.5 + 0.5 * math.log(2 * math.pi) + logstd).sum()
        return values.view(n, 1).expand(n, N).reshape(n * N, 1), logp.view(n * N, 1), ent

    # ---- compact-state path (include/dcc_b200.h "compact-state learner path") ------------------------
    def set_env_layout(self, pos_pois, m_energy=5.0):
        """Hands the learner handle the env's PoI table: from then on the *_state methods evaluate the first layer of both
        nets from the env's compact state (pos_vel (.., N, 4) float64, energy (.., M) uint8) instead of observation rows.
        Returns False (and changes nothing) if the layout is not the env's 2N + 2 + 5M / centralised-critic one."""
        poi = np.ascontiguousarray(np.asarray(pos_pois, dtype=np.float64)).reshape(-1, 2)
        rc = self.lib.dcc_mappo_set_env_layout(self._h, int(poi.shape[0]), poi.ctypes.data, float(m_energy))
        if rc == -5:
            return False
        _lib.check(rc, "dcc_mappo_set_env_layout")
        self.n_pois = int(poi.shape[0])
        return True

    def _as_state(self, pos_vel, energy):
        if not (isinstance(pos_vel, torch.Tensor) and pos_vel.is_cuda and pos_vel.dtype == torch.float64):
            pos_vel = torch.as_tensor(np.asarray(pos_vel), dtype=torch.float64).to(self.device)
        if not (isinstance(energy, torch.Tensor) and energy.is_cuda and energy.dtype == torch.uint8):
            energy = torch.as_tensor(np.asarray(energy), dtype=torch.uint8).to(self.device)
        pos_vel, energy = pos_vel.contiguous(), energy.contiguous()
        n = pos_vel.numel() // (self.n_agents * 4)
        if n * self.n_agents * 4 != pos_vel.numel() or energy.numel() != n * self.n_pois:
            raise ValueError("state shapes: pos_vel (.., N, 4) float64 and energy (.., M) uint8 with the same leading rows")
        return pos_vel, energy, n

    def get_actions_state(self, pos_vel, energy, deterministic=False, out_actions=None, out_logp=None, out_values=None):
        """get_actions on the compact state of one vec-env step.  Same outputs as get_actions."""
        pos_vel, energy, n = self._as_state(pos_vel, energy)
        N = self.n_agents
        actions = out_actions if out_actions is not None else self._out("actions", (n * N, 2))
        logp = out_logp if out_logp is not None else self._out("logp", (n * N,))
        values = out_values if out_values is not None else self._out("values", (n,))
        self._rng_offset += 1
        _lib.check(self.lib.dcc_mappo_act_state(self._h, self._ptr(self.actor.params), self._ptr(self.critic.params),
                                                self._ptr(pos_vel), self._ptr(energy), n, self.seed, self._rng_offset,
                                                1 if deterministic else 0, self._ptr(actions), self._ptr(logp),
                                                self._ptr(values), self._stream()), "dcc_mappo_act_state")
        v = values.view(n, 1).expand(n, N).reshape(n * N, 1)
        return v, actions.view(n * N, 2), logp.view(n * N, 1), None, None

    def get_values_state(self, pos_vel, energy, out_values=None):
        pos_vel, energy, n = self._as_state(pos_vel, energy)
        values = out_values if out_values is not None else self._out("values", (n,))
        _lib.check(self.lib.dcc_mappo_act_state(self._h, None, self._ptr(self.critic.params), self._ptr(pos_vel),
                                                self._ptr(energy), n, 0, 0, 0, None, None, self._ptr(values),
                                                self._stream()), "dcc_mappo_act_state(values)")
        return values.view(n, 1).expand(n, self.n_agents).reshape(n * self.n_agents, 1)

    def evaluate_actions_state(self, pos_vel, energy, action):
        """evaluate_actions on compact state rows: (values (B,1), action_log_probs (B,1), dist_entropy)."""
        pos_vel, energy, n = self._as_state(pos_vel, energy)
        N = self.n_agents
        action = torch.as_tensor(action, dtype=torch.float32, device=self.device).contiguous()
        logp = self._out("ev_logp", (n * N,))
        values = self._out("ev_values", (n,))
        _lib.check(self.lib.dcc_mappo_evaluate_state(self._h, self._ptr(self.actor.params), self._ptr(self.critic.params),
                                                     self._ptr(pos_vel), self._ptr(energy), self._ptr(action), n,
                                                     self._ptr(logp), self._ptr(values), None, self._stream()),
                   "dcc_mappo_evaluate_state")
        logstd = self.actor.view("act.action_out.logstd._bias")
        ent = (0.5 + 0.5 * math.log(2 * math.pi) + logstd).sum()
        return values.view(n, 1).expand(n, N).reshape(n * N, 1), logp.view(n * N, 1), ent

    def obs_from_state(self, pos_vel, energy, out=None):
        """Observation rows (rows, N, D) float32 regenerated from compact state rows (bit-identical to env.step's)."""
        pos_vel, energy, n = self._as_state(pos_vel, energy)
        obs = out if out is not None else torch.empty((n, self.n_agents, self.obs_dim), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.dcc_obs_from_state(self._h, self._ptr(pos_vel), self._ptr(energy), n, self._ptr(obs),
                                               self._stream()), "dcc_obs_from_state")
        return obs

    def act(self, obs, rnn_states_actor=None, masks=None, available_actions=None, deterministic=False):
        obs, n = self._as_obs(obs)
        actions = self._out("actions", (n * self.n_agents, 2))
        self._rng_offset += 1
        if self.recurrent_N:
            ha, _ = self._act_rnn(obs, n, rnn_states_actor, None, masks, 0, deterministic, actions, None, None, do_critic=False)
            return actions, ha
        _lib.check(self.lib.dcc_mappo_act(self._h, self._ptr(self.actor.params), None, self._ptr(obs), n, self.seed,
                                          self._rng_offset, 1 if deterministic else 0, self._ptr(actions), None, None,
                                          self._stream()), "dcc_mappo_act(actor)")
        return actions, rnn_states_actor

    # ---- checkpoints -----------------------------------------------------------------------------------------
    def state_dict(self):
        return {"actor": self.actor.state_dict(), "critic": self.critic.state_dict(),
                "actor_optimizer": {"step": self.actor.adam_step, "exp_avg": self.actor.adam_m.clone(),
                                    "exp_avg_sq": self.actor.adam_v.clone()},
                "critic_optimizer": {"step": self.critic.adam_step, "exp_avg": self.critic.adam_m.clone(),
                                     "exp_avg_sq": self.critic.adam_v.clone()}}

    def load_state_dict(self, sd):
        self.actor.load_state_dict(sd["actor"])
        self.critic.load_state_dict(sd["critic"])
        for net, key in ((self.actor, "actor_optimizer"), (self.critic, "critic_optimizer")):
            if key in sd:
                net.adam_step = int(sd[key]["step"])
                net.adam_m.copy_(sd[key]["exp_avg"])
                net.adam_v.copy_(sd[key]["exp_avg_sq"])

    def close(self):
        if getattr(self, "_h", None):
            self.lib.dcc_mappo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MAPPOTrainer:
    """algos/mappo.py:68-247.  `train(buffer)` runs the reference's update — advantage normalisation, ppo_epoch
    passes over the rollout (one minibatch = the whole rollout, or `num_mini_batch` random index lists per epoch,
    buffer/shared_buffer.py:219-279), ValueNorm update before every optimiser step, separate grad-norm clips, two
    Adam steps — as kernels; with a process group, gradients are all-reduced once per optimiser step."""

    def __init__(self, cfg, policy, agent_id=0, comm=None):
        self.policy = policy
        self.clip_param = cfg.clip_param
        self.ppo_epoch = int(cfg.ppo_epoch)
        self.num_mini_batch = int(cfg.num_mini_batch)
        self.value_loss_coef = cfg.value_loss_coef
        self.entropy_coef = cfg.entropy_coef
        self.max_grad_norm = cfg.max_grad_norm
        self.huber_delta = cfg.huber_delta
        self._use_valuenorm = bool(getattr(cfg, "use_valuenorm", True))
        # mappo.py:96-101 (PopArt is refused by check_supported): ValueNorm, or no normaliser at all
        self.value_normalizer = ValueNorm(1, device=policy.device) if self._use_valuenorm else None
        self.comm = comm if comm is not None else Comm()
        dev = policy.device
        n_upd = self.ppo_epoch * self.num_mini_batch
        self._stats4 = torch.zeros(4, dtype=torch.float64, device=dev)
        self._epoch_stats = torch.zeros((n_upd, 4), dtype=torch.float64, device=dev)
        self._gnorm_sq = torch.zeros((n_upd, 2), dtype=torch.float64, device=dev)
        self._mb_sums = torch.zeros((n_upd, 2), dtype=torch.float64, device=dev)
        self.permutation_fn = None   # tests inject the reference's permutations: fn(epoch, B) -> int64 tensor
        self.training = False
        # mappo.py:83-84,204-209: use_recurrent_policy wins over use_naive_recurrent_policy
        self._use_recurrent_policy = bool(getattr(cfg, "use_recurrent_policy", False))
        self._use_naive_recurrent = bool(getattr(cfg, "use_naive_recurrent_policy", False))
        self.data_chunk_length = int(getattr(cfg, "data_chunk_length", 10))

    def _permutation(self, epoch, n):
        """`torch.randperm(batch_size)` of feed_forward_generator (shared_buffer.py:238), drawn on the device."""
        if self.permutation_fn is not None:
            return torch.as_tensor(self.permutation_fn(epoch, n), dtype=torch.int64).to(self.policy.device).contiguous()
        return torch.randperm(n, device=self.policy.device)

    def _apply(self, u, update_actor):
        p, lib, ptr, s = self.policy, self.policy.lib, self.policy._ptr, self.policy._stream()
        self.comm.all_reduce_sum_(p.flat_grads)            # the one data-path collective
        for which, net, lr in ((0, p.actor, p.lr_actor_now), (1, p.critic, p.lr_critic_now)):
            if which == 0 and not update_actor:
                continue
            net.adam_step += 1
            _lib.check(lib.dcc_mappo_apply(p._h, which, ptr(net.params), ptr(net.grads), ptr(net.adam_m),
                                           ptr(net.adam_v), float(lr), net.adam_step,
                                           ptr(self._gnorm_sq[u, which:which + 1]), s), "dcc_mappo_apply")

    def train(self, buffer, update_actor=True):
        p, lib = self.policy, self.policy.lib
        # E = rows of the per-env arrays (values, returns, rewards): env instances, or (env, agent) pairs when the critic
        # is decentralised; N = agent rows per such row (p.n_agents is 1 in that case)
        T, E, N = buffer.episode_length, buffer.n_value_rows, p.n_agents
        s = p._stream()
        ptr = p._ptr
        vn = self.value_normalizer.state if self.value_normalizer is not None else None
        _lib.check(lib.dcc_mappo_train_begin(p._h, ptr(buffer.returns_te), ptr(buffer.values_te), ptr(vn), T, E,
                                             ptr(self._stats4), s), "dcc_mappo_train_begin")
        self.comm.all_reduce_sum_(self._stats4)
        rows_global = float(T) * float(E) * self.comm.world if getattr(buffer, "n_envs_global", None) is None \
            else float(T) * float(buffer.n_envs_global) * (E // buffer.n_rollout_threads)
        self._epoch_stats.zero_()
        self._gnorm_sq.zero_()
        nmb = self.num_mini_batch
        B_local = T * E * N
        mbs = B_local // nmb                               # mini_batch_size (shared_buffer.py:236); the tail is dropped
        mbs_global = float(mbs) * self.comm.world
        if p.recurrent_N:
            return self._train_recurrent(buffer, update_actor, rows_global, vn)
        for ep in range(self.ppo_epoch):
            if nmb == 1 and getattr(buffer, "compact", False):
                _lib.check(lib.dcc_mappo_epoch_grads_state(
                    p._h, ptr(p.actor.params), ptr(p.critic.params), ptr(p.actor.grads), ptr(p.critic.grads),
                    ptr(buffer.state_pv), ptr(buffer.state_en), ptr(buffer.actions), ptr(buffer.action_log_probs_ten),
                    ptr(buffer.values_te), ptr(buffer.returns_te), ptr(vn), ptr(self._stats4), rows_global, T, E,
                    ptr(self._epoch_stats[ep]), s), "dcc_mappo_epoch_grads_state")
                self._apply(ep, update_actor)
                continue
            if nmb == 1:
                _lib.check(lib.dcc_mappo_epoch_grads(
                    p._h, ptr(p.actor.params), ptr(p.critic.params), ptr(p.actor.grads), ptr(p.critic.grads),
                    ptr(buffer.obs), ptr(buffer.actions), ptr(buffer.action_log_probs_ten), ptr(buffer.values_te),
                    ptr(buffer.returns_te), ptr(vn), ptr(self._stats4), rows_global, T, E, ptr(self._epoch_stats[ep]), s),
                    "dcc_mappo_epoch_grads")
                self._apply(ep, update_actor)
                continue
            perm = self._permutation(ep, B_local)
            for i in range(nmb):
                u = ep * nmb + i
                idx = perm[i * mbs:(i + 1) * mbs]
                _lib.check(lib.dcc_mappo_minibatch_stats(p._h, ptr(buffer.returns_te), ptr(idx), mbs,
                                                         ptr(self._mb_sums[u]), s), "dcc_mappo_minibatch_stats")
                self.comm.all_reduce_sum_(self._mb_sums[u])
                if getattr(buffer, "compact", False):
                    _lib.check(lib.dcc_mappo_minibatch_grads_state(
                        p._h, ptr(p.actor.params), ptr(p.critic.params), ptr(p.actor.grads), ptr(p.critic.grads),
                        ptr(buffer.state_pv), ptr(buffer.state_en), ptr(buffer.actions), ptr(buffer.action_log_probs_ten),
                        ptr(buffer.values_te), ptr(buffer.returns_te), ptr(vn), ptr(self._stats4), rows_global, ptr(idx), mbs,
                        ptr(self._mb_sums[u]), mbs_global, ptr(self._epoch_stats[u]), s), "dcc_mappo_minibatch_grads_state")
                    self._apply(u, update_actor)
                    continue
                _lib.check(lib.dcc_mappo_minibatch_grads(
                    p._h, ptr(p.actor.params), ptr(p.critic.params), ptr(p.actor.grads), ptr(p.critic.grads),
                    ptr(buffer.obs), ptr(buffer.actions), ptr(buffer.action_log_probs_ten), ptr(buffer.values_te),
                    ptr(buffer.returns_te), ptr(vn), ptr(self._stats4), rows_global, ptr(idx), mbs,
                    ptr(self._mb_sums[u]), mbs_global, ptr(self._epoch_stats[u]), s), "dcc_mappo_minibatch_grads")
                self._apply(u, update_actor)
        return self._train_info(rows_global * N if nmb == 1 else mbs_global)

    def _train_recurrent(self, buffer, update_actor, rows_global, vn):
        """The update of a recurrent policy (mappo.py:203-209 -> shared_buffer.py:281-470).  Both generators cut the rollout,
        flattened in (env, agent, time) order, into chunks of L consecutive entries — L = data_chunk_length
        (recurrent_generator) or the whole episode (naive_recurrent_generator) — permute the chunks, split them into
        num_mini_batch minibatches (the tail is dropped) and feed each minibatch time-major, every chunk started from the
        hidden state the rollout stored in front of its first entry.  The chunk -> rollout-row index arithmetic happens
        here (a few tiny tensor ops per update); forward, loss and BPTT are dcc_mappo_seq_grads."""
        p, lib, ptr, s = self.policy, self.policy.lib, self.policy._ptr, self.policy._stream()
        T, E, N = buffer.episode_length, buffer.n_rollout_threads, p.n_agents
        B = T * E * N
        L = self.data_chunk_length if self._use_recurrent_policy else T
        n_chunks = B // L
        nmb = self.num_mini_batch
        mbs = n_chunks // nmb
        if mbs < 1:
            raise ValueError("fewer sequence chunks (%d) than PPO minibatches (%d)" % (n_chunks, nmb))
        P = int(lib.dcc_mappo_rnn_pass_seqs(p._h, L))
        if P < 1:
            raise _lib.DccError("sequence length %d exceeds the learner's activation chunk (%d rows): raise chunk_rows"
                                % (L, lib.dcc_mappo_chunk_rows(p._h)))
        rows_mb_global = float(mbs * L) * self.comm.world
        steps = torch.arange(L, device=p.device, dtype=torch.int64)[:, None]
        for ep in range(self.ppo_epoch):
            perm = self._permutation(ep, n_chunks)
            for i in range(nmb):
                u = ep * nmb + i
                ch = perm[i * mbs:(i + 1) * mbs]
                f = ch[None, :] * L + steps                                  # (L, chunks): entries in (env, agent, time) order
                ea, t = torch.div(f, T, rounding_mode="floor"), f % T
                rows = t * (E * N) + ea                                      # ea = e * N + a: rows of the (T, E, N) arrays
                # passes of P sequences, time-major inside a pass (the layout dcc_mappo_seq_grads reads)
                idx = torch.cat([rows[:, q:q + P].reshape(-1) for q in range(0, mbs, P)]).contiguous()
                _lib.check(lib.dcc_mappo_minibatch_stats(p._h, ptr(buffer.returns_te), ptr(idx), idx.numel(),
                                                         ptr(self._mb_sums[u]), s), "dcc_mappo_minibatch_stats")
                self.comm.all_reduce_sum_(self._mb_sums[u])
                _lib.check(lib.dcc_mappo_seq_grads(
                    p._h, ptr(p.actor.params), ptr(p.critic.params), ptr(p.actor.grads), ptr(p.critic.grads), ptr(buffer.obs),
                    ptr(buffer.rnn_a), ptr(buffer.rnn_c), ptr(buffer.masks_te), ptr(buffer.actions),
                    ptr(buffer.action_log_probs_ten), ptr(buffer.values_te), ptr(buffer.returns_te), ptr(vn), ptr(self._stats4),
                    rows_global, ptr(idx), mbs, L, ptr(self._mb_sums[u]), rows_mb_global, ptr(self._epoch_stats[u]), s),
                    "dcc_mappo_seq_grads")
                self._apply(u, update_actor)
        return self._train_info(rows_mb_global)

    def _train_info(self, B):
        """per-update sums -> the reference's train_info (means over updates; one device->host read per update).
        B = agent rows behind each update's loss mean."""
        nmb = self.num_mini_batch
        es = self._epoch_stats.clone()
        es[:, 3] = 0
        self.comm.all_reduce_sum_(es)
        es = es.cpu().numpy()
        ent = self._epoch_stats[:, 3].cpu().numpy()
        gn = np.sqrt(self._gnorm_sq.cpu().numpy())
        k = float(self.ppo_epoch * nmb)
        return {"value_loss": float(es[:, 1].sum() / B / k), "policy_loss": float(es[:, 0].sum() / B / k),
                "dist_entropy": float(ent.sum() / k), "actor_grad_norm": float(gn[:, 0].sum() / k),
                "critic_grad_norm": float(gn[:, 1].sum() / k), "ratio": float(es[:, 2].sum() / B / k)}

    def prep_training(self):
        self.training = True

    def prep_rollout(self):
        self.training = False

    def save_model(self, save_path):
        """Reference: pickle of the policy object (mappo.py:237-240).  Here: a pickle of plain CPU tensors keyed by
        the reference's state_dict names (+ Adam moments, the ValueNorm state and the sampling-stream position, which
        the reference drops), tagged with a format key."""
        sd = self.policy.state_dict()
        if self.value_normalizer is not None:
            sd["value_normalizer"] = self.value_normalizer.state_dict()
        sd["format"] = CHECKPOINT_FORMAT
        sd["rng"] = {"seed": int(self.policy.seed), "offset": int(self.policy._rng_offset)}
        cpu = _to_cpu(sd)
        with open(os.path.join(save_path, "agent.pkl"), "wb") as f:
            pickle.dump(cpu, f)

    def load_model(self, load_path):
        """Reads `agent.pkl` in either format: this build's dict of tensors, or a checkpoint written by the REFERENCE
        (`pickle.dump(self.policy, f)`, mappo.py:237-240 — the whole MAPPOPolicy object with its R_Actor / R_Critic
        modules and Adam optimisers).  The file is read with a restricted unpickler (torch / numpy / collections
        only; the reference's own classes become inert stand-ins), never with a bare pickle.load."""
        path = os.path.join(load_path, "agent.pkl") if os.path.isdir(load_path) else load_path
        sd = load_checkpoint(path)
        self.policy.load_state_dict(sd)
        if "value_normalizer" in sd and self.value_normalizer is not None:
            self.value_normalizer.load_state_dict(sd["value_normalizer"])
        if "rng" in sd:
            self.policy.seed, self.policy._rng_offset = int(sd["rng"]["seed"]), int(sd["rng"]["offset"])


CHECKPOINT_FORMAT = "dcc_b200.agent.v2"


class _InertObject:
    """Stand-in for a class of the reference (algos.*, utils.*, gym.*) met while unpickling its agent.pkl: absorbs the
    pickled state, runs none of the original code."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)
        else:
            self.__dict__["_state"] = state


class _CheckpointUnpickler(pickle.Unpickler):
    _ALLOWED_ROOTS = ("torch", "numpy", "collections", "builtins", "copyreg", "_codecs", "argparse")
    _BUILTINS_OK = {"dict", "list", "tuple", "set", "frozenset", "int", "float", "bool", "str", "bytes", "bytearray",
                    "complex", "slice", "range", "object", "getattr"}

    def find_class(self, module, name):
        root = module.split(".")[0]
        if root == "builtins":
            if name not in self._BUILTINS_OK:
                raise pickle.UnpicklingError("refusing builtins.%s in a checkpoint" % name)
            return super().find_class(module, name)
        if root in self._ALLOWED_ROOTS:
            return super().find_class(module, name)
        return type(name, (_InertObject,), {"__module__": module})


def _module_tensors(mod, prefix=""):
    """state_dict of a (real or inert) nn.Module tree in torch's order: parameters, persistent buffers, submodules."""
    out = OrderedDict()
    d = mod.__dict__
    for k, v in (d.get("_parameters") or {}).items():
        if v is not None:
            out[prefix + k] = v.detach()
    skip = d.get("_non_persistent_buffers_set") or set()
    for k, v in (d.get("_buffers") or {}).items():
        if v is not None and k not in skip:
            out[prefix + k] = v.detach()
    for k, sub in (d.get("_modules") or {}).items():
        if sub is not None:
            out.update(_module_tensors(sub, prefix + k + "."))
    return out


def load_checkpoint(path):
    """-> {"actor": state_dict, "critic": state_dict, [optimizers, value_normalizer, rng]} from an agent.pkl written by
    this build OR by the reference (see MAPPOTrainer.load_model)."""
    with open(path, "rb") as f:
        obj = _CheckpointUnpickler(f).load()
    if isinstance(obj, dict):
        if "actor" not in obj or "critic" not in obj:
            raise ValueError("%s: a dict without 'actor' / 'critic' entries is not an agent checkpoint" % path)
        return obj
    if hasattr(obj, "actor") and hasattr(obj, "critic"):          # the reference's pickled MAPPOPolicy object
        sd = {"actor": _module_tensors(obj.actor), "critic": _module_tensors(obj.critic), "format": "reference.policy.pickle"}
        if not sd["actor"] or not sd["critic"]:
            raise ValueError("%s: no parameters found in the pickled policy object" % path)
        return sd
    raise ValueError("%s holds a %s, neither this build's checkpoint dict nor a reference MAPPOPolicy object"
                     % (path, type(obj).__name__))


def _to_cpu(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu()
    if isinstance(x, dict):
        return {k: _to_cpu(v) for k, v in x.items()}
    return x
