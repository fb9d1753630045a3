"""make_env(cfg) — the reference's env factory boundary (envs/make_env.py:8-49), CUDA-backed.

Same call, same config keys (config/env_config/dcc.yaml): `env_file`, `num_agents`, `num_pois`, `max_ep_len`,
`r_cover`, `r_comm`, `comm_r_scale`, `comm_force_scale`, `n_rollout_threads`, `seed`.  Where the reference
returns DummyVecEnv / SubprocVecEnv (one Python env, or one OS process per env), this returns ONE
CudaVecEnv stepping all `n_rollout_threads` instances in a single launch.

Optional new keys (defaults preserve the shipped behaviour):
  reference_compat (True)   the shipped scenario never forwards comm_r_scale/comm_force_scale to the world
                            (scenarios/coverage.py:34); False passes them through ("connectivity active")
  pos_pois / pos_pois_path  PoI layout (M,2) or an .npy file; DEFAULT: the reference's own layout,
                            scenarios/pos_pois.npy[0:M] (scenarios/coverage.py:15-17, shipped as envs/data/pos_pois.npy)
  poi_layout                "reference" (default) or "synthetic" / "synthetic:<seed>" = uniform(-1,1) layout, the
                            reference's commented alternative (coverage.py:18) used by the throughput benchmarks
  per_env_layouts (False)   every env instance gets its own uniform(-1,1) PoI layout (generator seeded with
                            `poi_seed`, default 0; rank-offset in multi-GPU jobs) — the reference's commented
                            `np.random.uniform(-1, 1)` alternative (coverage.py:71) made per-instance;
                            pos_pois may also be given as an (E, M, 2) array
  numpy_compat (False)      numpy in/out with the reference's dtypes instead of CUDA tensors
  device (0)                CUDA device index
`cfg.seed` is accepted and ignored exactly like the reference (env.seed only seeds numpy, which the env
never draws from: environment.py:124-125, core.py:79).
"""
import numpy as np

from .cuda_vec_env import CudaVecEnv


def make_env(cfg, **kwargs):
    if kwargs is not None:
        for k, v in kwargs.items():
            setattr(cfg, k, v)
    if "uav_dcc" not in cfg.env_file:
        raise NotImplementedError("env_file: %s not found" % cfg.env_file)
    pos_pois = getattr(cfg, "pos_pois", None)
    path = getattr(cfg, "pos_pois_path", None)
    if pos_pois is None and path:
        pos_pois = np.load(path)[0:cfg.num_pois, :]
    if pos_pois is None:
        pos_pois = getattr(cfg, "poi_layout", None)          # None -> the reference table (CudaVecEnv default)
    per_env = None
    if pos_pois is not None and not isinstance(pos_pois, str) and np.asarray(pos_pois).ndim == 3:
        per_env, pos_pois = np.asarray(pos_pois, dtype=np.float64), np.asarray(pos_pois, dtype=np.float64)[0]
    elif getattr(cfg, "per_env_layouts", False):
        rng = np.random.default_rng(int(getattr(cfg, "poi_seed", 0)) + 7919 * int(getattr(cfg, "env_rank", 0)))
        per_env = rng.uniform(-1.0, 1.0, (int(cfg.n_rollout_threads), int(cfg.num_pois), 2))
    env = CudaVecEnv(
        n_envs=cfg.n_rollout_threads,
        num_agents=cfg.num_agents,
        num_pois=cfg.num_pois,
        max_ep_len=cfg.max_ep_len,
        r_cover=cfg.r_cover,
        r_comm=cfg.r_comm,
        comm_r_scale=cfg.comm_r_scale,
        comm_force_scale=cfg.comm_force_scale,
        reference_compat=bool(getattr(cfg, "reference_compat", True)),
        pos_pois=pos_pois,
        device=int(getattr(cfg, "device", 0) or 0),
        numpy_compat=bool(getattr(cfg, "numpy_compat", False)),
        want_connectivity=bool(getattr(cfg, "want_connectivity", False)),
    )
    if per_env is not None:
        env.set_poi_layouts(per_env)
    return env
