"""SharedReplayBuffer — the reference's rollout storage (buffer/shared_buffer.py:14-279), resident in HBM.

Same constructor and method names (`insert`, `compute_returns`, `after_update`) and the same logical arrays, but:
  * tensors live on the GPU and the env / policy kernels write straight into them (no numpy, no per-step copies of
    obs: `obs[t+1]` IS the env kernel's output buffer for step t);
  * `share_obs` is never materialised: the centralised observation of env e is its N obs rows concatenated
    (learner.py:219-220, 270-271), i.e. `obs` viewed as (T+1, E, N*D);
  * quantities that are identical for the N agents of an env (reward, mask, value prediction, return — shared reward
    environment.py:106-108, identical critic input) are stored once per env: (T(+1), E).  The reference-shaped
    (T(+1), E, N, 1) arrays are exposed as expanded views (`rewards`, `masks`, `value_preds`, `returns`);
  * the rnn_states arrays (never written with MLP policies), bad_masks and active_masks (all ones) are not stored:
    `rnn_states`, `rnn_states_critic`, `bad_masks`, `active_masks` are zero-cost expanded views with the reference's
    shapes; `available_actions` is None as in the reference (Box action space).  Recurrent policies
    (use_recurrent_policy / use_naive_recurrent_policy) DO store them: `rnn_a` (T+1, E, N, recurrent_N, H) for the actor
    and `rnn_c` (T+1, E, recurrent_N, H) for the critic — one per env, because the N agent rows of an env feed the
    critic identical inputs, masks and therefore hidden states; `rnn_states_critic` is the expanded view;
GAE (`compute_returns`, shared_buffer.py:199-208) is one kernel, `dcc_mappo_gae`.

Two storage modes for the observations:
  * materialised (`compact=False`): `obs` is the full (T+1, E, N, D) float32 tensor — 107 GB at 8 UAV / 64 PoI /
    65 536 envs / T = 150 of the 180 GB HBM3e.  Needed when the caller hands observations in (`insert`), for
    `num_mini_batch > 1`, per-env PoI layouts and the decentralised critic.
  * compact (`compact=True`, what `Learner` picks whenever the path allows it): the rollout keeps the env's compact
    state per step — `state_pv` (T+1, E, N, 4) float64 and `state_en` (T+1, E, M) uint8, 320 B per env step at 8/64,
    3.2 GB instead of 107 GB — and the learner kernels evaluate the first layer from it directly (exact algebra,
    csrc/dcc_compact.cuh).  `obs` then is a lazy reference-shaped view: `buffer.obs[t]` regenerates step t's
    (E, N, D) rows bit-identically to what the env kernel would have written.
"""
import ctypes as C

import torch

from .. import _lib


class RegeneratedObs:
    """Reference-shaped lazy view of a compact rollout's observations: indexing a step (or a slice of steps)
    regenerates the float32 rows from the stored compact state (dcc_obs_from_state)."""

    def __init__(self, buf):
        self._buf = buf
        T1, E, N, _ = buf.state_pv.shape
        self.shape = (T1, E, N, buf.obs_dim)
        self.dtype = torch.float32
        self.device = buf.device

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, idx):
        b = self._buf
        T1, E, N, D = self.shape
        if isinstance(idx, int):
            t = idx % T1
            return b.regenerate_obs(t, t + 1)[0]
        if isinstance(idx, slice):
            lo, hi, st = idx.indices(T1)
            if st != 1:
                raise IndexError("RegeneratedObs supports contiguous step ranges only")
            return b.regenerate_obs(lo, hi)
        raise IndexError("RegeneratedObs is indexed by a step or a contiguous range of steps")

    def materialize(self):
        return self[0:self.shape[0]]


class SharedReplayBuffer(object):
    def __init__(self, cfg, obs_space, cent_obs_space, act_space, device=None, compact=False, n_pois=None):
        self.lib = _lib.load()
        self.episode_length = int(cfg.max_ep_len)
        self.n_rollout_threads = int(cfg.n_rollout_threads)
        self.gamma, self.gae_lambda = float(cfg.gamma), float(cfg.gae_lambda)
        self.num_agents = int(cfg.num_agents)
        self._use_gae = bool(getattr(cfg, "use_gae", True))
        self._use_valuenorm = bool(getattr(cfg, "use_valuenorm", True))
        self.obs_dim = int(obs_space.shape[0])
        self.act_dim = int(act_space.shape[0])
        self.device = torch.device("cuda", int(getattr(cfg, "device", 0) or 0)) if device is None else torch.device(device)
        T, E, N, D = self.episode_length, self.n_rollout_threads, self.num_agents, self.obs_dim
        # centralised critic (share_obs = the env's N rows concatenated): one value / return / reward row per env.
        # use_centralized_V: false (cent_obs_space is obs_space, learner.py:43-46): one per (env, agent), and the learner
        # kernels see E*N pseudo-envs with one agent each.
        self.centralized = int(cent_obs_space.shape[0]) != self.obs_dim or N == 1
        self.n_value_rows = E if self.centralized else E * N
        self.agents_per_value_row = N if self.centralized else 1
        V = self.n_value_rows
        kw = dict(dtype=torch.float32, device=self.device)
        self.compact = bool(compact)
        self._policy = None
        if self.compact:
            if not self.centralized or n_pois is None:
                raise ValueError("compact rollout storage needs the centralised critic and the env's PoI count")
            self.n_pois = int(n_pois)
            if D != 4 + 2 * (N - 1) + 5 * self.n_pois:
                raise ValueError("obs_dim %d is not the env's 2N + 2 + 5M layout" % D)
            self.state_pv = torch.zeros((T + 1, E, N, 4), dtype=torch.float64, device=self.device)
            self.state_en = torch.zeros((T + 1, E, self.n_pois), dtype=torch.uint8, device=self.device)
            self.obs = RegeneratedObs(self)
        else:
            self.obs = torch.zeros((T + 1, E, N, D), **kw)
        self.actions = torch.zeros((T, E, N, self.act_dim), **kw)
        self.action_log_probs_ten = torch.zeros((T, E, N), **kw)   # one column (the reference stores 2 equal ones)
        self.values_te = torch.zeros((T + 1, V), **kw)
        self.returns_te = torch.zeros((T + 1, V), **kw)
        self.rewards_te = torch.zeros((T, V), **kw)
        self.masks_te = torch.ones((T + 1, V), **kw)
        self.step = 0
        self.recurrent_N = int(getattr(cfg, "recurrent_N", 1))
        self.hidden_size = int(getattr(cfg, "algo_hidden_size", 256))
        from ..utils.config import is_recurrent
        self.recurrent = is_recurrent(cfg)
        if self.recurrent:
            if self.compact or not self.centralized:
                raise ValueError("recurrent policies use materialised observations and the centralised critic")
            self.rnn_a = torch.zeros((T + 1, E, N, self.recurrent_N, self.hidden_size), **kw)
            self.rnn_c = torch.zeros((T + 1, E, self.recurrent_N, self.hidden_size), **kw)
        self.available_actions = None
        self._zero = torch.zeros(1, **kw)
        self._one = torch.ones(1, **kw)

    # ---- reference-shaped views (no copies) ---------------------------------------------------------------
    @property
    def rnn_states(self):
        """(T+1, E, N, recurrent_N, hidden) zeros (shared_buffer.py:44-48): MLP policies never write them."""
        if self.recurrent:
            return self.rnn_a
        T1, E, N, _ = self.obs.shape
        return self._zero.view(1, 1, 1, 1, 1).expand(T1, E, N, self.recurrent_N, self.hidden_size)

    @property
    def rnn_states_critic(self):
        if self.recurrent:
            T1, E = self.rnn_c.shape[:2]
            return self.rnn_c[:, :, None].expand(T1, E, self.num_agents, self.recurrent_N, self.hidden_size)
        return self.rnn_states

    @property
    def bad_masks(self):
        """(T+1, E, N, 1) ones: the reference never stores bad masks (learner.py:272-276)."""
        T1, E, N, _ = self.obs.shape
        return self._one.view(1, 1, 1, 1).expand(T1, E, N, 1)

    @property
    def active_masks(self):
        return self.bad_masks

    @property
    def share_obs(self):
        T1, E, N, D = self.obs.shape
        if not self.centralized:
            return self.obs
        obs = self.obs.materialize() if self.compact else self.obs
        return obs.view(T1, E, 1, N * D).expand(T1, E, N, N * D)

    # ---- compact mode ------------------------------------------------------------------------------------
    def attach_policy(self, policy):
        """The learner handle that knows the env's PoI table (MAPPOPolicy.set_env_layout): regenerates `obs[t]`."""
        self._policy = policy

    def regenerate_obs(self, t0, t1):
        """(t1 - t0, E, N, D) float32 observation rows of steps [t0, t1) rebuilt from the compact state."""
        if self._policy is None:
            raise RuntimeError("compact buffer: attach_policy(policy) first (the PoI table lives in the learner handle)")
        return self._policy.obs_from_state(self.state_pv[t0:t1], self.state_en[t0:t1]).view(
            t1 - t0, self.n_rollout_threads, self.num_agents, self.obs_dim)

    def _per_agent(self, x):
        if not self.centralized:
            return x.view(x.shape[0], self.n_rollout_threads, self.num_agents, 1)
        return x[:, :, None, None].expand(x.shape[0], x.shape[1], self.num_agents, 1)

    @property
    def value_preds(self):
        return self._per_agent(self.values_te)

    @property
    def returns(self):
        return self._per_agent(self.returns_te)

    @property
    def rewards(self):
        return self._per_agent(self.rewards_te)

    @property
    def masks(self):
        return self._per_agent(self.masks_te)

    @property
    def action_log_probs(self):
        return self.action_log_probs_ten[..., None].expand(*self.action_log_probs_ten.shape, self.act_dim)

    # ---- writers -----------------------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def insert(self, share_obs, obs, rnn_states_actor, rnn_states_critic, actions, action_log_probs, value_preds,
               rewards, masks, bad_masks=None, active_masks=None, available_actions=None):
        """shared_buffer.py:72-105 with CUDA tensors.  share_obs / rnn states are accepted and ignored.  Arguments
        that already ARE the destination slice (the zero-copy rollout) are not copied again."""
        t, E, N = self.step, self.n_rollout_threads, self.num_agents

        def put(dst, src):
            if src is None:
                return
            src = torch.as_tensor(src, dtype=torch.float32, device=self.device)
            if src.data_ptr() != dst.data_ptr():
                dst.copy_(src.reshape(dst.shape))

        def per_env(x):   # (E,N,1) / (E*N,1) / (E,) -> one entry per value row
            x = torch.as_tensor(x, dtype=torch.float32, device=self.device)
            if x.numel() == self.n_value_rows:
                return x.reshape(self.n_value_rows)
            return x.reshape(E, N, -1)[:, 0, 0]

        if self.compact:
            if obs is not None:
                raise ValueError("compact rollout storage takes the env state (state_pv / state_en), not observation rows")
        else:
            put(self.obs[t + 1], obs)
        if self.recurrent:      # shared_buffer.py:88-89: the states AFTER step t (already zeroed where the episode ended)
            put(self.rnn_a[t + 1], rnn_states_actor)
            if rnn_states_critic is not None:
                hc = torch.as_tensor(rnn_states_critic, dtype=torch.float32, device=self.device)
                if hc.numel() == self.rnn_c[t + 1].numel() * N:      # the reference's per-agent layout: N identical rows per env
                    hc = hc.reshape(E, N, self.recurrent_N, self.hidden_size)[:, 0]
                put(self.rnn_c[t + 1], hc)
        put(self.actions[t], actions)
        alp = torch.as_tensor(action_log_probs, dtype=torch.float32, device=self.device)
        put(self.action_log_probs_ten[t], alp if alp.numel() == E * N else alp.reshape(E, N, -1)[..., 0])
        put(self.values_te[t], per_env(value_preds))
        put(self.rewards_te[t], per_env(rewards))
        put(self.masks_te[t + 1], per_env(masks))
        self.step = (self.step + 1) % self.episode_length

    def insert_env_step(self, rew_en, done_en):
        """Zero-copy rollout step: obs[t+1], actions[t], log-probs[t] and values[t] were written in place by the env
        and policy kernels; this stores the per-env reward and mask = 1 - done (learner.py:266-267) — one kernel.
        Recurrent policies: the new hidden states were written to slot t+1 by the policy kernels; those of finished
        episodes are zeroed here (learner.py:258-265)."""
        t = self.step
        _lib.check(self.lib.dcc_rollout_insert(C.c_void_p(rew_en.data_ptr()), C.c_void_p(done_en.data_ptr()),
                                               self.n_value_rows, self.agents_per_value_row,
                                               C.c_void_p(self.rewards_te[t].data_ptr()),
                                               C.c_void_p(self.masks_te[t + 1].data_ptr()), self._stream()),
                   "dcc_rollout_insert")
        if self.recurrent:
            m = self.masks_te[t + 1]
            self.rnn_a[t + 1].mul_(m.view(-1, 1, 1, 1))
            self.rnn_c[t + 1].mul_(m.view(-1, 1, 1))
        self.step = (self.step + 1) % self.episode_length

    def compute_returns(self, next_value, value_normalizer, policy=None):
        """shared_buffer.py:154-212: GAE or plain discounted returns (`use_gae`), with or without a value normaliser
        (`value_normalizer` None = use_valuenorm false); `use_proper_time_limits` changes nothing in the reference
        because it never stores bad_masks (they stay 1, learner.py:272-276).  The branch taken is the learner
        handle's cfg.  next_value: (E,) / (E*N,1) / (E,N,1), or None if the
        bootstrap value already sits in values_te[T]."""
        T, E, N = self.episode_length, self.n_value_rows, self.num_agents
        if next_value is not None:
            nv = torch.as_tensor(next_value, dtype=torch.float32, device=self.device)
            nv = nv if nv.numel() == E else nv.reshape(self.n_rollout_threads, N, -1)[:, 0, 0]
            if nv.data_ptr() != self.values_te[T].data_ptr():
                self.values_te[T].copy_(nv.reshape(E))
        h = policy._h if policy is not None else self._gae_handle()
        p = lambda t: C.c_void_p(t.data_ptr())   # noqa: E731
        _lib.check(self.lib.dcc_mappo_gae(h, p(self.rewards_te), p(self.values_te), p(self.masks_te),
                                          p(value_normalizer.state) if value_normalizer is not None else None, T, E,
                                          p(self.returns_te), self._stream()),
                   "dcc_mappo_gae")

    def _gae_handle(self):
        # a minimal learner handle just for its gamma / lambda (when the caller has no policy at hand)
        if getattr(self, "_h", None) is None:
            mc = _lib.MappoCfg()
            _lib.check(self.lib.dcc_mappo_cfg_default(C.byref(mc)), "dcc_mappo_cfg_default")
            mc.n_agents, mc.obs_dim, mc.hidden, mc.chunk_rows = 1, 1, 1, 64
            mc.gamma, mc.gae_lambda = self.gamma, self.gae_lambda
            mc.use_gae, mc.use_valuenorm = int(self._use_gae), int(self._use_valuenorm)
            h = C.c_void_p()
            _lib.check(self.lib.dcc_mappo_create(C.byref(mc), self.device.index, C.byref(h)), "dcc_mappo_create")
            self._h = h
        return self._h

    def after_update(self):
        """shared_buffer.py:142-152: the last step becomes the first of the next rollout."""
        if self.compact:
            self.state_pv[0].copy_(self.state_pv[-1])
            self.state_en[0].copy_(self.state_en[-1])
        else:
            self.obs[0].copy_(self.obs[-1])
        self.masks_te[0].copy_(self.masks_te[-1])
        if self.recurrent:      # shared_buffer.py:146-147: the hidden states carry over into the next rollout as well
            self.rnn_a[0].copy_(self.rnn_a[-1])
            self.rnn_c[0].copy_(self.rnn_c[-1])

    def close(self):
        if getattr(self, "_h", None) is not None:
            self.lib.dcc_mappo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
