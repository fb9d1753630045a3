"""train.py — the reference's entry point (train.py:10-29): `python -m dcc_b200.train <gpu_id> [config_dir]`.

Loads the reference's three YAML files from `config_dir` (default ./config, the reference's layout) — unchanged —
merges them env < algo < expt, applies the two overrides the reference hard-codes (log_wandb False, save_model True)
and runs `Learner(cfg).train()` on the CUDA path.  Under torchrun the env axis is sharded across the ranks.
"""
import os
import sys

import torch

from .learner import Learner
from .parallel import init_from_env
from .utils.config import load_config


def main(argv=None):
    argv = sys.argv[1:] if argv is None else list(argv)
    gpu_id = int(argv[0]) if argv else 0
    config_dir = argv[1] if len(argv) > 1 else ("./config" if os.path.isdir("./config") else None)
    overrides = {}
    for kv in argv[2:]:          # key=value overrides, e.g. n_rollout_threads=65536 num_agents=8 num_pois=64
        k, v = kv.split("=", 1)
        try:
            v = int(v)
        except ValueError:
            try:
                v = float(v)
            except ValueError:
                v = {"true": True, "false": False, "none": None}.get(v.lower(), v)
        overrides[k] = v
    comm = init_from_env()
    if comm.world > 1:
        gpu_id = int(os.environ.get("LOCAL_RANK", gpu_id))
    cfg = load_config(config_dir, **overrides)
    cfg.device = gpu_id
    print("cuda is available: ", torch.cuda.is_available())
    if not torch.cuda.is_available():
        raise SystemExit("the B200 path needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(gpu_id)
    os.makedirs(cfg.main_save_path, exist_ok=True)
    cfg.log_wandb = False       # train.py:25
    cfg.save_model = True       # train.py:26
    learner = Learner(cfg, comm=comm)
    learner.train()


if __name__ == "__main__":
    main()
