"""CudaVecEnv — the vec-env object `make_env(cfg)` returns, backed by one CUDA launch per step.

Mirrors the surface Learner uses on the reference's SubprocVecEnv / DummyVecEnv (envs/wrappers.py:133-261):
`reset()`, `step(actions)`, `close()`, `.observation_space`, `.share_observation_space`, `.action_space`
(lists of length N), `.n_agents`, `.n_envs` / `.num_envs`.

Two I/O modes, same kernel:
  * tensor mode (default): `step(actions)` takes a CUDA float32 tensor (E,N,2) and returns CUDA tensors
    (obs (E,N,D) f32, rewards (E,N,1) f32, dones (E,N) bool, infos) with no host round trip;
  * numpy_compat=True: takes/returns numpy arrays with the reference's exact shapes and dtypes
    (obs float64 (E,N,D), rewards float64 (E,N,1), dones bool (E,N), infos = tuple of dicts), going through
    the host-buffer C entry point `dcc_env_step_host` — what the parity tests and the e2e bench use.
All paths raise DccError if the CUDA library is missing; there is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np
import torch

from .. import _lib
from .spaces import Box


def synthetic_pois(n_pois, seed=0):
    """Synthetic PoI layout: uniform in [-1,1]^2 (the reference's own commented alternative,
    scenarios/coverage.py:18).  What the throughput benchmarks use (BASELINE.json: "synthetic PoI layouts")."""
    return np.random.default_rng(seed).uniform(-1.0, 1.0, (n_pois, 2))


REFERENCE_POIS_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "pos_pois.npy")


def reference_pois(n_pois):
    """The reference's own PoI layout: the first `n_pois` rows of its scenarios/pos_pois.npy (scenarios/coverage.py:15-17;
    1000 x 2 float64, shipped here as a data table).  This is the DEFAULT layout, so that the unchanged YAMLs train on
    the same map as the reference."""
    table = np.load(REFERENCE_POIS_PATH)
    if n_pois > table.shape[0]:
        raise ValueError("the reference PoI table has %d rows; num_pois = %d needs an explicit pos_pois / "
                         "poi_layout: synthetic" % (table.shape[0], n_pois))
    return np.ascontiguousarray(table[0:n_pois, :], dtype=np.float64)


def resolve_pois(pos_pois, n_pois):
    """pos_pois: None / "reference" -> reference_pois; "synthetic" (or "synthetic:<seed>") -> synthetic_pois; else an array."""
    if pos_pois is None or (isinstance(pos_pois, str) and pos_pois == "reference"):
        return reference_pois(n_pois)
    if isinstance(pos_pois, str):
        if pos_pois.startswith("synthetic"):
            seed = int(pos_pois.split(":")[1]) if ":" in pos_pois else 0
            return synthetic_pois(n_pois, seed)
        raise ValueError("unknown PoI layout %r (use 'reference', 'synthetic[:seed]' or an (M,2) array)" % pos_pois)
    return pos_pois


class CoverageInfos:
    """Lazy stand-in for the reference's per-env info dicts (uav_dcc.py:46-49): `infos[e]["coverage_rate"]`.
    `.coverage_rate` is the (E,) float32 CUDA tensor; indexing materialises it on the host once."""

    def __init__(self, coverage_rate):
        self.coverage_rate = coverage_rate
        self._host = None

    def __len__(self):
        return int(self.coverage_rate.shape[0])

    def __getitem__(self, e):
        if self._host is None:
            cr = self.coverage_rate
            self._host = cr.detach().cpu().numpy() if isinstance(cr, torch.Tensor) else np.asarray(cr)
        return {"coverage_rate": float(self._host[e]), "n": []}

    def __iter__(self):
        return (self[e] for e in range(len(self)))


class CudaVecEnv:
    def __init__(self, n_envs, num_agents=4, num_pois=20, max_ep_len=150, r_cover=0.2, r_comm=0.4,
                 comm_r_scale=0.95, comm_force_scale=0.0, reference_compat=True, pos_pois=None, device=0,
                 numpy_compat=False, want_connectivity=False):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.DccError("CudaVecEnv needs a CUDA device (sm_100); there is no CPU fallback")
        self.device = torch.device("cuda", int(device))
        self.n_envs = self.num_envs = int(n_envs)
        self.n_agents = int(num_agents)
        self.n_pois = int(num_pois)
        self.max_ep_len = int(max_ep_len)
        self.numpy_compat = bool(numpy_compat)
        self.obs_dim = int(self.lib.dcc_env_obs_dim(self.n_agents, self.n_pois))
        E, N, D = self.n_envs, self.n_agents, self.obs_dim

        cfg = _lib.EnvCfg()
        _lib.check(self.lib.dcc_env_cfg_default(C.byref(cfg)), "dcc_env_cfg_default")
        cfg.n_envs, cfg.n_agents, cfg.n_pois, cfg.max_ep_len = E, N, self.n_pois, self.max_ep_len
        cfg.r_cover, cfg.r_comm = float(r_cover), float(r_comm)
        cfg.comm_r_scale, cfg.comm_force_scale = float(comm_r_scale), float(comm_force_scale)
        cfg.reference_compat = 1 if reference_compat else 0
        self.cfg = cfg
        pos_pois = resolve_pois(pos_pois, self.n_pois)   # default: the reference's pos_pois.npy[0:M]
        self.pos_pois = np.ascontiguousarray(np.asarray(pos_pois, dtype=np.float64)[: self.n_pois]).reshape(self.n_pois, 2)
        h = C.c_void_p()
        _lib.check(self.lib.dcc_env_create(C.byref(cfg), self.pos_pois.ctypes.data, self.device.index, C.byref(h)),
                   "dcc_env_create")
        self._h = h

        # spaces exactly as DCEnv declares them (uav_dcc.py:38-43, environment.py:52,72)
        self.observation_space = [Box(-np.inf, np.inf, (D,)) for _ in range(N)]
        self.share_observation_space = [Box(-np.inf, np.inf, (N * D,)) for _ in range(N)]
        self.action_space = [Box(-1.0, 1.0, (2,)) for _ in range(N)]

        self.want_connectivity = bool(want_connectivity)
        with torch.cuda.device(self.device):
            self.obs = torch.empty((E, N, D), dtype=torch.float32, device=self.device)
            self.rewards = torch.empty((E, N, 1), dtype=torch.float32, device=self.device)
            self.dones_u8 = torch.empty((E, N), dtype=torch.uint8, device=self.device)
            self.coverage_rate = torch.empty((E,), dtype=torch.float32, device=self.device)
            self.connect_bits = torch.zeros((E,), dtype=torch.uint8, device=self.device)
            if self.want_connectivity:
                self.adj = torch.zeros((E, N), dtype=torch.int32, device=self.device)
                self.adj_s = torch.zeros((E, N), dtype=torch.int32, device=self.device)
            else:
                self.adj = self.adj_s = None
        self._host = None
        self.closed = False
        self.pos_pois_per_env = None

    # ---- helpers -------------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _ptr(t):
        return None if t is None else C.c_void_p(t.data_ptr())

    @property
    def share_obs(self):
        """(E, N*D) view of the observation buffer = the critic's centralised input (learner.py:219-220)."""
        return self.obs.view(self.n_envs, -1)

    def launch_count(self):
        return int(self.lib.dcc_env_launch_count(self._h))

    def set_poi_layouts(self, pos_pois_per_env):
        """Per-env PoI layouts: array (E, M, 2), or None to return to the shared layout.  Takes effect from the next
        reset / step (include/dcc_b200.h: dcc_env_set_poi_layouts)."""
        if pos_pois_per_env is None:
            self.pos_pois_per_env = None
            _lib.check(self.lib.dcc_env_set_poi_layouts(self._h, None, self._stream()), "dcc_env_set_poi_layouts")
            return
        pp = np.ascontiguousarray(np.asarray(pos_pois_per_env, dtype=np.float64))
        if pp.shape != (self.n_envs, self.n_pois, 2):
            raise ValueError("per-env PoI layouts must have shape %s, got %s" % ((self.n_envs, self.n_pois, 2), pp.shape))
        self.pos_pois_per_env = pp
        _lib.check(self.lib.dcc_env_set_poi_layouts(self._h, pp.ctypes.data, self._stream()), "dcc_env_set_poi_layouts")

    def use_specialized(self, enable=True):
        """Toggle the compile-time specialised kernel (4/20, 8/64, 16/256).  Returns True if it is in use."""
        return bool(self.lib.dcc_env_use_specialized(self._h, 1 if enable else 0))

    def set_launch(self, warps_per_cta, ctas=0):
        _lib.check(self.lib.dcc_env_set_launch(self._h, int(warps_per_cta), int(ctas)), "dcc_env_set_launch")

    # ---- reference surface ---------------------------------------------------------------------------
    def _obs_out(self, out_obs):
        if out_obs is None:
            return self.obs
        if not (isinstance(out_obs, torch.Tensor) and out_obs.is_cuda and out_obs.dtype == torch.float32 and
                out_obs.is_contiguous() and out_obs.numel() == self.obs.numel()):
            raise ValueError("out_obs must be a contiguous CUDA float32 tensor with E*N*D elements")
        return out_obs

    def snapshot_state_into(self, pos_vel_out, energy_out):
        """Device copy of the live compact state into rollout storage: pos_vel_out (E, N, 4) float64, energy_out (E, M)
        uint8 CUDA tensors (SharedReplayBuffer.state_pv[t] / state_en[t] of a compact rollout)."""
        if not (pos_vel_out.is_cuda and pos_vel_out.dtype == torch.float64 and pos_vel_out.is_contiguous() and
                pos_vel_out.numel() == self.n_envs * self.n_agents * 4 and energy_out.is_cuda and
                energy_out.dtype == torch.uint8 and energy_out.is_contiguous() and
                energy_out.numel() == self.n_envs * self.n_pois):
            raise ValueError("snapshot_state_into needs contiguous CUDA tensors (E,N,4) float64 and (E,M) uint8")
        _lib.check(self.lib.dcc_env_snapshot_state(self._h, self._ptr(pos_vel_out), self._ptr(energy_out), self._stream()),
                   "dcc_env_snapshot_state")

    def reset(self, out_obs=None, write_obs=True):
        """out_obs (tensor mode): write the observations there instead of the env's own buffer — the rollout
        storage passes `buffer.obs[0]` so nothing is copied afterwards.  write_obs=False (compact rollouts): only the
        state is reset, no observation row is written."""
        if self.numpy_compat:
            hb = self._host_buffers()
            _lib.check(self.lib.dcc_env_reset_host(self._h, hb["obs"].ctypes.data, self._stream()), "dcc_env_reset_host")
            return hb["obs"].astype(np.float64)
        obs = self._obs_out(out_obs) if write_obs else None
        _lib.check(self.lib.dcc_env_reset(self._h, self._ptr(obs), self._stream()), "dcc_env_reset")
        return obs

    def step(self, actions, out_obs=None, write_obs=True):
        """write_obs=False (compact rollouts, tensor mode): the step updates the state, rewards, dones and infos but
        writes no observation rows (returned obs is None) — the learner reads the compact state instead."""
        if self.numpy_compat:
            return self._step_numpy(actions)
        if not (isinstance(actions, torch.Tensor) and actions.is_cuda and actions.dtype == torch.float32):
            raise TypeError("tensor mode expects a CUDA float32 tensor of shape (E,N,2); use numpy_compat=True for numpy")
        if actions.numel() != self.n_envs * self.n_agents * 2:
            raise ValueError("actions has %d elements, expected E*N*2 = %d" % (actions.numel(), self.n_envs * self.n_agents * 2))
        actions = actions.contiguous()
        obs = self._obs_out(out_obs) if write_obs else None
        _lib.check(self.lib.dcc_env_step(self._h, self._ptr(actions), self._ptr(obs), self._ptr(self.rewards),
                                         self._ptr(self.dones_u8), self._ptr(self.coverage_rate),
                                         self._ptr(self.connect_bits), self._ptr(self.adj), self._ptr(self.adj_s),
                                         self._stream()), "dcc_env_step")
        return (None if obs is None else obs.view(self.n_envs, self.n_agents, self.obs_dim)), self.rewards, \
            self.dones_u8.view(torch.bool), CoverageInfos(self.coverage_rate)

    def step_async(self, actions):
        self._pending = actions

    def step_wait(self):
        a, self._pending = self._pending, None
        return self.step(a)

    def close(self):
        if not self.closed and getattr(self, "_h", None):
            self.lib.dcc_env_destroy(self._h)
            self._h = None
        self.closed = True

    def enable_connectivity_outputs(self):
        """Allocate the per-UAV adjacency outputs (filled by every following step)."""
        if not self.want_connectivity:
            with torch.cuda.device(self.device):
                self.adj = torch.zeros((self.n_envs, self.n_agents), dtype=torch.int32, device=self.device)
                self.adj_s = torch.zeros((self.n_envs, self.n_agents), dtype=torch.int32, device=self.device)
            self.want_connectivity = True

    def snapshot(self, max_envs=None):
        """Compact state of the first `max_envs` env instances on the host (synchronises): dict with pos_vel
        (e,N,4) float64, energy (e,M) uint8, connect_bits (e,) uint8 [bit0 connect, bit1 connect_], adj (e,N) uint32
        bitmask rows of the comm graph (None until enable_connectivity_outputs), coverage_rate (e,), reward (e,)."""
        e = self.n_envs if max_envs is None else min(int(max_envs), self.n_envs)
        pv, en = self.get_state()
        return dict(pos_vel=pv[:e], energy=en[:e], connect_bits=self.connect_bits[:e].cpu().numpy(),
                    adj=None if self.adj is None else self.adj[:e].cpu().numpy().view(np.uint32),
                    coverage_rate=self.coverage_rate[:e].cpu().numpy(), reward=self.rewards[:e, 0, 0].cpu().numpy())

    def render(self, mode="human", max_envs=1, size=350):
        """Headless replacement of the pyglet viewer (environment.py:209-330): mode "rgb_array" returns, like the
        reference's vec-env, a list over envs of one-element lists holding an (size,size,3) uint8 frame
        (`frame[0][0]` is what learner.py:201 keeps); "human" has no window to draw into and returns None.
        Visualisation only — it copies the compact state to the host."""
        if mode != "rgb_array":
            return None
        from .headless_render import rasterize
        st = self.snapshot(max_envs)
        frames = []
        for e in range(st["pos_vel"].shape[0]):
            pos = st["pos_vel"][e, :, :2]
            if st["adj"] is not None:
                rows = st["adj"][e]
            else:   # comm graph of the frame from the positions: d < r_a + r_b (CoverageWorld.py:78-80)
                d = np.sqrt(((pos[:, None] - pos[None]) ** 2).sum(-1)) + np.eye(self.n_agents) * 1e5
                rows = ((d < 2.0 * self.cfg.r_comm) * (1 << np.arange(self.n_agents))[None]).sum(1).astype(np.uint32)
            poi = self.pos_pois if self.pos_pois_per_env is None else self.pos_pois_per_env[e]
            frames.append([rasterize(pos, poi, st["energy"][e], rows, self.cfg.r_cover, self.cfg.r_comm,
                                     self.cfg.m_energy, size=size)])
        return frames

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- host-buffer path (numpy_compat / e2e) ---------------------------------------------------------
    def _host_buffers(self):
        if self._host is None:
            E, N, D = self.n_envs, self.n_agents, self.obs_dim
            self._host = dict(
                actions=torch.empty((E, N, 2), dtype=torch.float32).pin_memory().numpy(),
                obs=torch.empty((E, N, D), dtype=torch.float32).pin_memory().numpy(),
                rew=torch.empty((E, N), dtype=torch.float32).pin_memory().numpy(),
                done=torch.empty((E, N), dtype=torch.uint8).pin_memory().numpy(),
                cov=torch.empty((E,), dtype=torch.float32).pin_memory().numpy())
        return self._host

    def step_host(self, actions_np):
        """Host-buffer step through `dcc_env_step_host` (pinned staging; H2D + kernel + D2H + sync).
        Returns views of the pinned float32 result buffers (overwritten by the next call)."""
        hb = self._host_buffers()
        if actions_np is not hb["actions"]:
            np.copyto(hb["actions"], np.asarray(actions_np, dtype=np.float32).reshape(hb["actions"].shape))
        _lib.check(self.lib.dcc_env_step_host(self._h, hb["actions"].ctypes.data, hb["obs"].ctypes.data,
                                              hb["rew"].ctypes.data, hb["done"].ctypes.data, hb["cov"].ctypes.data,
                                              self._stream()), "dcc_env_step_host")
        return hb["obs"], hb["rew"], hb["done"], hb["cov"]

    def _step_numpy(self, actions):
        a = np.asarray(actions)
        if a.shape != (self.n_envs, self.n_agents, 2):
            raise ValueError("actions shape %s != %s" % (a.shape, (self.n_envs, self.n_agents, 2)))
        obs, rew, done, cov = self.step_host(a)  # the caller's array is NOT scaled in place (environment.py:186-190 does)
        infos = tuple({"coverage_rate": float(c), "n": []} for c in cov)
        return (obs.astype(np.float64), rew.astype(np.float64).reshape(self.n_envs, self.n_agents, 1),
                done.astype(bool), infos)

    # ---- parity / checkpoint access to the compact state -------------------------------------------------
    def get_state(self):
        pv = np.empty((self.n_envs, self.n_agents, 4), dtype=np.float64)
        en = np.empty((self.n_envs, self.n_pois), dtype=np.uint8)
        _lib.check(self.lib.dcc_env_get_state(self._h, pv.ctypes.data, en.ctypes.data, self._stream()), "dcc_env_get_state")
        return pv, en

    def set_state(self, pos_vel, energy):
        pv = np.ascontiguousarray(pos_vel, dtype=np.float64).reshape(self.n_envs, self.n_agents, 4)
        en = np.ascontiguousarray(energy, dtype=np.uint8).reshape(self.n_envs, self.n_pois)
        _lib.check(self.lib.dcc_env_set_state(self._h, pv.ctypes.data, en.ctypes.data, self._stream()), "dcc_env_set_state")
