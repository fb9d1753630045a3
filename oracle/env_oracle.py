"""ctypes front end of the CPU env oracle (oracle/dcc_env_oracle.c).  TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, `__graft_entry__.smoke()` and bench.py's cpu_baseline / `--impl reference` leg import this
module; the product package never does.  See the C file's header for what is restated (reference
file:line) and how it is pinned (tests/golden/env_*.npz generated from the unmodified reference).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdcc_oracle.so")


class OracleCfg(C.Structure):
    _fields_ = [("n_agents", C.c_int32), ("n_pois", C.c_int32)] + [
        (k, C.c_double) for k in ("r_cover", "r_comm", "comm_r_scale", "contact_force", "contact_margin", "dt",
                                  "damping", "max_speed", "sensitivity", "m_energy", "rew_cover", "rew_done",
                                  "rew_out")]


def build(force=False):
    """Compile the oracle with gcc (no OpenMP dependency; POSIX threads only)."""
    src = os.path.join(HERE, "dcc_env_oracle.c")
    if (not force and os.path.exists(LIB_PATH)
            and os.path.getmtime(LIB_PATH) >= os.path.getmtime(src)):
        return LIB_PATH
    cmd = ["gcc", "-O2", "-fPIC", "-std=gnu11", "-ffp-contract=off", "-fno-fast-math", "-shared",
           "-o", LIB_PATH, src, "-lm", "-lpthread"]
    subprocess.run(cmd, check=True)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.dcc_oracle_step.restype = C.c_int
        _lib.dcc_oracle_reset.restype = C.c_int
        _lib.dcc_oracle_step_layouts.restype = C.c_int
        _lib.dcc_oracle_reset_layouts.restype = C.c_int
        _lib.dcc_oracle_max_threads.restype = C.c_int
    return _lib


def make_cfg(n_agents, n_pois, r_cover=0.2, r_comm=0.4, comm_r_scale=0.95, contact_force=0.0):
    """`comm_r_scale` / `contact_force` are the WORLD's values (shipped: 0.9 / 0.0; generalised:
    cfg.comm_r_scale / 100*cfg.comm_force_scale) — SURVEY.md Appendix C.1."""
    return OracleCfg(n_agents, n_pois, r_cover, r_comm, comm_r_scale, contact_force, 1e-3, 0.1, 0.25, 0.5, 5.0,
                     5.0, 75.0, 1500.0, -100.0)


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


class OracleEnv:
    """E independent env instances stepped on the CPU (float64 state), reference semantics."""

    def __init__(self, n_envs, n_agents, n_pois, poi_xy, r_cover=0.2, r_comm=0.4, comm_r_scale=0.95,
                 contact_force=0.0, n_threads=1):
        self.E, self.N, self.M = n_envs, n_agents, n_pois
        self.D = 4 + 2 * (n_agents - 1) + 5 * n_pois
        self.cfg = make_cfg(n_agents, n_pois, r_cover, r_comm, comm_r_scale, contact_force)
        poi_xy = np.ascontiguousarray(poi_xy, dtype=np.float64)
        # (M, 2): one table for all envs (the reference's pos_pois.npy); (E, M, 2): every env has its own layout
        self.per_env = poi_xy.ndim == 3
        self.poi = poi_xy.reshape((n_envs, n_pois, 2) if self.per_env else (n_pois, 2))
        self.poi_stride = 2 * n_pois if self.per_env else 0
        self.pos_vel = np.zeros((n_envs, n_agents, 4), dtype=np.float64)
        self.energy = np.zeros((n_envs, n_pois), dtype=np.uint8)
        self.n_threads = n_threads

    def reset(self, want_obs=True):
        obs = np.empty((self.E, self.N, self.D), dtype=np.float32) if want_obs else None
        rc = lib().dcc_oracle_reset_layouts(C.byref(self.cfg), self.E, _p(self.poi, C.c_double), int(self.poi_stride),
                                            _p(self.pos_vel, C.c_double), _p(self.energy, C.c_uint8), _p(obs, C.c_float))
        assert rc == 0
        return obs

    def set_state(self, pos_vel, energy):
        self.pos_vel[...] = np.asarray(pos_vel, dtype=np.float64).reshape(self.pos_vel.shape)
        self.energy[...] = np.asarray(energy, dtype=np.uint8).reshape(self.energy.shape)

    def step(self, actions, want_obs=True, want_aux=True):
        a = np.ascontiguousarray(actions, dtype=np.float32).reshape(self.E, self.N, 2)
        E, N, M = self.E, self.N, self.M
        out = dict(
            obs=np.empty((E, N, self.D), dtype=np.float32) if want_obs else None,
            reward=np.empty(E, dtype=np.float64), done=np.empty(E, dtype=np.uint8),
            coverage_rate=np.empty(E, dtype=np.float64), connect_bits=np.empty(E, dtype=np.uint8),
            adj=np.empty((E, N), dtype=np.uint32) if want_aux else None,
            adj_=np.empty((E, N), dtype=np.uint32) if want_aux else None,
            pos_vel_pre=np.empty((E, N, 4), dtype=np.float64) if want_aux else None,
            energy_pre=np.empty((E, M), dtype=np.uint8) if want_aux else None)
        rc = lib().dcc_oracle_step_layouts(
            C.byref(self.cfg), E, _p(self.poi, C.c_double), int(self.poi_stride), _p(a, C.c_float),
            _p(self.pos_vel, C.c_double),
            _p(self.energy, C.c_uint8), _p(out["obs"], C.c_float), _p(out["reward"], C.c_double),
            _p(out["done"], C.c_uint8), _p(out["coverage_rate"], C.c_double), _p(out["connect_bits"], C.c_uint8),
            _p(out["adj"], C.c_uint32), _p(out["adj_"], C.c_uint32), _p(out["pos_vel_pre"], C.c_double),
            _p(out["energy_pre"], C.c_uint8), int(self.n_threads))
        assert rc == 0
        out["connect"] = (out["connect_bits"] & 1).astype(bool)
        out["connect_"] = ((out["connect_bits"] >> 1) & 1).astype(bool)
        out["done"] = out["done"].astype(bool)
        out["pos_vel"] = self.pos_vel.copy()
        out["energy"] = self.energy.copy()
        return out


def max_threads():
    return int(lib().dcc_oracle_max_threads())
