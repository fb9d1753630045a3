"""Config loading for the re-hosted trainer: the reference's three YAML files drive it unchanged.

The reference merges config/env_config/dcc.yaml < config/algo_config/mappo.yaml < config/expt.yaml with OmegaConf,
later files winning (train.py:12-19), and turns the result into an argparse.Namespace (learner.py:23).  OmegaConf is
not a dependency here: PyYAML with the float resolver OmegaConf applies (`5e-4` must load as a float).

`DEFAULTS` holds the effective merged values of the shipped files so the trainer also runs with no
config directory at all; new optional keys of this build sit at the end and default to the shipped behaviour.
"""
import os
import re
from argparse import Namespace

DEFAULTS = dict(
    # env (dcc.yaml)
    env_file="mpe.uav_dcc", env_class="DCEnv", scenario_name="coverage", num_agents=4, num_pois=20, max_ep_len=150,
    r_cover=0.2, r_comm=0.4, comm_r_scale=0.95, comm_force_scale=0.0, save_name="uav_dcc", ppo_epoch=15,
    n_rollout_threads=16,
    # algo (mappo.yaml)
    algo_file="mappo", n_training_threads=32, n_eval_rollout_threads=1, n_render_rollout_threads=1,
    use_obs_instead_of_state=False, share_policy=False, use_centralized_V=True, use_stacked_frames=False,
    stacked_frames=1, algo_hidden_size=256, layer_N=1, use_ReLU=True, use_popart=False, use_valuenorm=True,
    use_feature_normalization=True, use_orthogonal=True, gain=0.01, use_recurrent_policy=False,
    use_naive_recurrent_policy=False, recurrent_N=1, data_chunk_length=10, actor_lr=5e-4, critic_lr=5e-4,
    opti_eps=1e-5, weight_decay=0, use_clipped_value_loss=True, clip_param=0.2, num_mini_batch=1, entropy_coef=0.01,
    value_loss_coef=1, use_max_grad_norm=True, max_grad_norm=10.0, use_gae=True, gamma=0.99, gae_lambda=0.95,
    use_proper_time_limits=False, use_huber_loss=True, use_value_active_masks=True, use_policy_active_masks=True,
    huber_delta=10.0, use_linear_lr_decay=True, use_render=False, render_episodes=5, ifi=0.1,
    # experiment (expt.yaml)
    seed=0, n_iters=200, eval_interval=10, render_interval=200, save_gifs=True, save_interval=50, log_wandb=True,
    log_interval=1, save_model=True, load_model=False, load_buffer_path=None, main_save_path="results/",
    load_model_path="results/dcc/xxxx_xxxx_sdx/models_xxx.pt", hidden_sizes_mlp=[64], lr=5e-4,
    # new optional keys of the B200 build
    reference_compat=True, pos_pois_path=None, poi_layout=None, compact_rollout=None, per_env_layouts=False, poi_seed=0, device=0, chunk_rows=0, gemm_backend=0,
)

_FLOAT = re.compile(r"^[-+]?(\d+\.?\d*|\.\d+)([eE][-+]?\d+)?$")


def _loader():
    import yaml

    class Loader(yaml.SafeLoader):
        pass

    Loader.add_implicit_resolver("tag:yaml.org,2002:float", _FLOAT, list("-+0123456789."))
    return yaml, Loader


def load_yaml(path):
    yaml, Loader = _loader()
    with open(path) as f:
        data = yaml.load(f, Loader=Loader)
    return dict(data or {})


def load_config(config_dir=None, **overrides):
    """Merged config as a Namespace.  config_dir = a directory laid out like the reference's `config/`
    (env_config/dcc.yaml, algo_config/mappo.yaml, expt.yaml); missing files fall back to DEFAULTS."""
    cfg = dict(DEFAULTS)
    if config_dir:
        for rel in ("env_config/dcc.yaml", "algo_config/mappo.yaml", "expt.yaml"):   # later wins (train.py:12-19)
            p = os.path.join(config_dir, rel)
            if os.path.exists(p):
                cfg.update(load_yaml(p))
    cfg.update(overrides)
    if cfg.get("load_buffer_path") == "None":
        cfg["load_buffer_path"] = None
    return Namespace(**cfg)


def is_recurrent(cfg):
    """r_actor_critic.py:36,102: either flag puts the RNNLayer into both nets."""
    return bool(getattr(cfg, "use_recurrent_policy", False) or getattr(cfg, "use_naive_recurrent_policy", False))


def check_supported(cfg):
    """Branches of mappo.yaml the B200 path implements: everything the shipped configuration enables plus the
    update-path switches (use_huber_loss, use_clipped_value_loss, use_max_grad_norm, use_valuenorm, use_gae,
    use_proper_time_limits, weight_decay, num_mini_batch, use_linear_lr_decay, the *_active_masks flags — no-ops
    in the reference, whose active masks are all ones) and the network switches use_ReLU (tanh trunk),
    use_feature_normalization, use_orthogonal (xavier init), use_centralized_V (per-agent critic), layer_N 1..3.
    Recurrent policies (use_recurrent_policy: chunked BPTT over data_chunk_length steps; use_naive_recurrent_policy: whole
    episodes; recurrent_N 1..4 GRU layers) run with the centralised critic on materialised observation rows.
    Refused loudly instead of silently computing something else: use_popart, which the unmodified reference itself
    cannot run (PopArt.update assigns a tensor to an nn.Parameter attribute and raises TypeError on the first update,
    popart.py:61)."""
    bad = []
    if is_recurrent(cfg):
        if not 1 <= int(getattr(cfg, "recurrent_N", 1)) <= 4:
            bad.append("recurrent_N outside 1..4")
        if not getattr(cfg, "use_centralized_V", True):
            bad.append("recurrent policies with use_centralized_V: false")
        if getattr(cfg, "use_recurrent_policy", False) and int(getattr(cfg, "data_chunk_length", 10)) < 1:
            bad.append("data_chunk_length < 1")
    if getattr(cfg, "use_popart", False):
        bad.append("use_popart")
    # use_stacked_frames / stacked_frames, use_obs_instead_of_state, share_policy, use_render, render_episodes, ifi:
    # read (mlp.py:39, learner.py:40) or merely listed in mappo.yaml but never acted on by the reference -> accepted, no-ops
    if int(getattr(cfg, "num_mini_batch", 1)) < 1:
        bad.append("num_mini_batch < 1")
    if not 1 <= int(getattr(cfg, "layer_N", 1)) <= 3:
        bad.append("layer_N outside 1..3")
    if bad:
        raise NotImplementedError("not supported by the B200 hot path: " + ", ".join(bad))
