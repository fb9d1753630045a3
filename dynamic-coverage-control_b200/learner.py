"""Learner — the reference's orchestrator (learner.py:21-322) re-hosted on the CUDA env + learner kernels.

Same methods (`train`, `rollout`, `warmup`, `collect`, `insert`, `compute`, `rl_update`, `log`, `save_model`,
`load_model`) and the same loop; what disappears is the numpy glue between them: the env kernel writes step t's
observations straight into `buffer.obs[t+1]`, the policy kernels read them there and write actions / log-probs /
values straight into the buffer, and `insert` is one tiny kernel (reward, mask).  Render rollouts
(`render_interval`, `n_render_rollout_threads` = 1) run headless: instead of a pyglet window they record the compact
state per step (`models_<iter>_traj.npz`: positions, PoI energy, connect bits, adjacency) and, with `save_gifs`,
a GIF rasterised by envs/headless_render.py.  `rollout_info` carries one key more than the reference's:
`connect_rate`, the fraction of (env, step) pairs whose UAV comm graph was connected (`world.connect`,
CoverageWorld.py:92 — computed by the reference every step but never surfaced; asset/cc.png plots it).

Multi-GPU (torchrun, one process per GPU): `n_rollout_threads` is the GLOBAL env count, sharded contiguously across
ranks (parallel.shard_envs); parameters are replicated (rank 0's initial weights are broadcast); the only data-path
collective is the per-epoch gradient all-reduce inside MAPPOTrainer.train.
"""
import copy
import datetime
import json
import os
import time
from argparse import Namespace

import torch

from .algos.mappo import MAPPOPolicy, MAPPOTrainer
from .buffer.shared_buffer import SharedReplayBuffer
from .envs.make_env import make_env
from .parallel import Comm, shard_envs


def seed_everything(seed):
    """utils/util.py:7-12."""
    import random
    import numpy as np
    random.seed(seed)
    torch.random.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    np.random.seed(seed)


class Learner:
    def __init__(self, cfg, comm=None):
        self.cfg = cfg if isinstance(cfg, Namespace) else Namespace(**dict(cfg))
        cfg = self.cfg
        seed_everything(cfg.seed)
        self.comm = comm if comm is not None else Comm()
        self.n_envs_global = int(cfg.n_rollout_threads)
        lo, hi = shard_envs(self.n_envs_global, self.comm.world, self.comm.rank)
        local = copy.copy(cfg)
        local.n_rollout_threads = hi - lo
        local.env_rank = self.comm.rank            # per-env synthetic PoI layouts differ across ranks (make_env)
        self.local_cfg = local

        # 1. env
        self.train_envs = make_env(cfg=copy.copy(local))
        self.n_agents = cfg.num_agents
        self.max_ep_len = cfg.max_ep_len
        self.obs_dim_n = [self.train_envs.observation_space[i].shape[0] for i in range(self.n_agents)]
        self.action_dim_n = [self.train_envs.action_space[i].shape[0] for i in range(self.n_agents)]

        # 2. rl agent (one shared policy for all agents, learner.py:48-57)
        # learner.py:43-46: a decentralised critic (use_centralized_V: false) takes the agent's own observation
        self.share_observation_space = self.train_envs.share_observation_space[0] if getattr(cfg, "use_centralized_V", True) \
            else self.train_envs.observation_space[0]
        self.policy = MAPPOPolicy(cfg, self.train_envs.observation_space[0], self.share_observation_space,
                                  self.train_envs.action_space[0], device=self.train_envs.device)
        self.trainer = MAPPOTrainer(cfg=cfg, policy=self.policy, comm=self.comm)
        if self.comm.world > 1:   # replicas start from rank 0's weights
            self.comm.broadcast_(self.policy.actor.params)
            self.comm.broadcast_(self.policy.critic.params)
            self.policy.seed = int(cfg.seed) + 7919 * self.comm.rank   # independent exploration noise per shard

        # 3. rollout storage.  Compact whenever the path allows it: the rollout keeps the env's compact
        # state (320 B per env step at 8/64 instead of 10.8 KB of observation rows) and the learner kernels evaluate
        # the first layer from it (also for num_mini_batch > 1: rows are gathered from the state through the permutation).
        # `compact_rollout: false` forces the materialised (T+1, E, N, D) observation tensor; per-env PoI layouts and a
        # decentralised critic need it and select it automatically.
        want = getattr(cfg, "compact_rollout", None)
        can = (getattr(cfg, "use_centralized_V", True) and not self.policy.recurrent_N and
               self.train_envs.pos_pois_per_env is None and not getattr(cfg, "numpy_compat", False))
        if want and not can:
            raise NotImplementedError("compact_rollout needs an MLP policy, use_centralized_V and a shared PoI layout")
        self.compact = bool(can if want is None else want)
        if self.compact:
            self.compact = self.policy.set_env_layout(self.train_envs.pos_pois, self.train_envs.cfg.m_energy)
        self.rl_buffer = self._make_buffer(local)
        self.rl_buffer.n_envs_global = self.n_envs_global
        if cfg.n_eval_rollout_threads > 0:
            test_cfg = copy.copy(cfg)
            test_cfg.n_rollout_threads = cfg.n_eval_rollout_threads
            self.test_envs = make_env(test_cfg)
            self.test_buffer = self._make_buffer(test_cfg)
        self.render_interval = int(getattr(cfg, "render_interval", 0) or 0)
        self.save_gifs = bool(getattr(cfg, "save_gifs", False))
        self.n_render = int(getattr(cfg, "n_render_rollout_threads", 0) or 0) if self.comm.rank == 0 else 0
        if self.n_render > 0:
            assert self.n_render == 1, "n_render_rollout_threads must be 1 (learner.py:82)"
            render_cfg = copy.copy(cfg)
            render_cfg.n_rollout_threads = self.n_render
            self.render_envs = make_env(render_cfg)
            self.render_envs.enable_connectivity_outputs()
            self.render_buffer = self._make_buffer(render_cfg)
        self.last_trajectory = None

        # 4. train-loop parameters
        self.use_linear_lr_decay = cfg.use_linear_lr_decay
        self.n_iters = cfg.n_iters
        self.eval_interval = cfg.eval_interval
        self.log_interval = cfg.log_interval
        self.is_save_model = bool(cfg.save_model) and self.comm.rank == 0
        self.save_interval = cfg.save_interval
        if cfg.load_model:
            self.load_model(cfg.load_model_path)
        self.expt_name = datetime.datetime.now().strftime("%m%d_%H%M_") + "sd{}".format(cfg.seed)
        if self.is_save_model:
            self.output_path = str(os.path.join(cfg.main_save_path, cfg.save_name, self.expt_name))
            os.makedirs(self.output_path, exist_ok=True)
            with open(os.path.join(self.output_path, "config.json"), "w") as f:
                json.dump({k: v for k, v in vars(cfg).items() if _jsonable(v)}, f, indent=4)
        self._start_time = time.time()
        self._check_time = time.time()
        self.agent_steps = 0

    def _make_buffer(self, cfg):
        buf = SharedReplayBuffer(cfg, self.train_envs.observation_space[0], self.share_observation_space,
                                 self.train_envs.action_space[0], device=self.train_envs.device, compact=self.compact,
                                 n_pois=self.train_envs.n_pois)
        if self.compact:
            buf.attach_policy(self.policy)
        return buf

    def train(self):
        self.warmup(self.rl_buffer, self.train_envs)
        for iter_ in range(1, self.n_iters + 1):
            if self.use_linear_lr_decay:
                self.trainer.policy.lr_decay(iter_, self.n_iters)
            rollout_info = self.rollout(self.rl_buffer, self.train_envs)
            rl_train_info = self.rl_update()
            if self.cfg.n_eval_rollout_threads > 0 and iter_ % self.eval_interval == 0:
                test_rollout_info = self.rollout(self.test_buffer, self.test_envs)
            else:
                test_rollout_info = {}
            if self.n_render > 0 and self.render_interval > 0 and iter_ % self.render_interval == 0:
                self.rollout(self.render_buffer, self.render_envs, is_render=True, iter_=iter_)
            if iter_ % self.log_interval == 0 and self.comm.rank == 0:
                self.log(iter_=iter_, rollout_info=rollout_info, rl_train_info=rl_train_info,
                         test_rollout_info=test_rollout_info)
            if self.is_save_model and (iter_ % self.save_interval == 0):
                save_path = os.path.join(self.output_path, "models_%d.pt" % iter_)
                os.makedirs(save_path, exist_ok=True)
                self.save_model(save_path)
                print("model saved in %s" % save_path)
        self.train_envs.close()
        if self.cfg.n_eval_rollout_threads > 0:
            self.test_envs.close()
        if self.n_render > 0:
            self.render_envs.close()

    # ---- collect ---------------------------------------------------------------------------------------------
    def rollout(self, r_buffer, r_envs, is_render=False, iter_=0):
        """learner.py:178-214.  Every rollout starts from a reset; returns {"reward", "coverage_rate"} as the
        reference does, plus "connect_rate".  is_render=True records the trajectory of the (single) render env
        instead of drawing it (see the module docstring)."""
        self.warmup(r_buffer, r_envs)
        E = r_buffer.n_rollout_threads
        rew_sum = torch.zeros((), dtype=torch.float32, device=r_buffer.device)
        sr = torch.zeros(E, dtype=torch.float32, device=r_buffer.device)
        conn = torch.zeros((), dtype=torch.float32, device=r_buffer.device)
        rec = None
        if is_render:
            from .envs.headless_render import TrajectoryRecorder
            rec = TrajectoryRecorder(r_envs.pos_pois, r_envs.cfg.r_cover, r_envs.cfg.r_comm)
        for cur_step in range(self.max_ep_len):
            actions = self.collect(cur_step, r_buffer)
            if r_buffer.compact:     # no observation rows: the step's compact state goes to slot t+1 of the rollout
                obs, rewards, dones, infos = r_envs.step(actions, write_obs=False)
                r_envs.snapshot_state_into(r_buffer.state_pv[cur_step + 1], r_buffer.state_en[cur_step + 1])
            else:
                obs, rewards, dones, infos = r_envs.step(actions, out_obs=r_buffer.obs[cur_step + 1])
            self.insert((obs, rewards, dones, infos), r_buffer, r_envs)
            rew_sum += rewards.mean()
            torch.maximum(sr, infos.coverage_rate, out=sr)
            conn += (r_envs.connect_bits & 1).float().mean()
            if rec is not None:
                st = r_envs.snapshot()
                rec.add(st["pos_vel"], st["energy"], st["connect_bits"], st["adj"], st["coverage_rate"], st["reward"])
        self.compute(r_buffer)
        if not is_render:
            self.agent_steps += self.max_ep_len * E * self.n_agents
        out = torch.stack([rew_sum, sr.mean(), conn / self.max_ep_len]).tolist()   # the rollout's only device->host read
        if rec is not None:
            self.last_trajectory = rec
            if self.is_save_model:
                rec.save(os.path.join(self.output_path, "models_%d_traj.npz" % iter_))
                if self.save_gifs:
                    rec.save_gif(os.path.join(self.output_path, "models_%d.gif" % iter_))
        return {"reward": out[0], "coverage_rate": out[1], "connect_rate": out[2]}

    def warmup(self, r_buffer, r_envs):
        if r_buffer.compact:
            r_envs.reset(write_obs=False)
            r_envs.snapshot_state_into(r_buffer.state_pv[0], r_buffer.state_en[0])
        else:
            r_envs.reset(out_obs=r_buffer.obs[0])

    def collect(self, cur_step, r_buffer):
        """learner.py:227-252: policy forward on step `cur_step`'s observations; actions, log-probs and values land
        in the buffer slices directly.  Returns the action tensor (E, N, 2) for the env."""
        self.trainer.prep_rollout()
        outs = dict(out_actions=r_buffer.actions[cur_step], out_logp=r_buffer.action_log_probs_ten[cur_step],
                    out_values=r_buffer.values_te[cur_step])
        if r_buffer.compact:
            self.trainer.policy.get_actions_state(r_buffer.state_pv[cur_step], r_buffer.state_en[cur_step], **outs)
        elif r_buffer.recurrent:     # learner.py:231-238: the step's stored hidden states and masks go in, slot t+1 receives the new ones
            self.trainer.policy.get_actions(None, r_buffer.obs[cur_step], r_buffer.rnn_a[cur_step], r_buffer.rnn_c[cur_step],
                                            r_buffer.masks_te[cur_step], out_rnn_actor=r_buffer.rnn_a[cur_step + 1],
                                            out_rnn_critic=r_buffer.rnn_c[cur_step + 1], **outs)
        else:
            self.trainer.policy.get_actions(None, r_buffer.obs[cur_step], **outs)
        return r_buffer.actions[cur_step]

    def insert(self, data, r_buffer, r_envs=None):
        """learner.py:254-276."""
        obs, rewards, dones, infos = data
        r_buffer.insert_env_step(rewards, dones.view(torch.uint8))

    def compute(self, r_buffer):
        """learner.py:278-287: bootstrap value from the last observations, then GAE."""
        self.trainer.prep_rollout()
        T = r_buffer.episode_length
        if r_buffer.compact:
            self.trainer.policy.get_values_state(r_buffer.state_pv[T], r_buffer.state_en[T], out_values=r_buffer.values_te[T])
        elif r_buffer.recurrent:
            self.trainer.policy.get_values(r_buffer.obs[T], r_buffer.rnn_c[T], r_buffer.masks_te[T], out_values=r_buffer.values_te[T])
        else:
            self.trainer.policy.get_values(r_buffer.obs[T], out_values=r_buffer.values_te[T])
        r_buffer.compute_returns(None, self.trainer.value_normalizer, policy=self.policy)

    # ---- update ------------------------------------------------------------------------------------------------
    def rl_update(self):
        self.trainer.prep_training()
        update_info = self.trainer.train(buffer=self.rl_buffer, update_actor=True)
        self.rl_buffer.after_update()
        return update_info

    # ---- log / save / load -------------------------------------------------------------------------------------
    def log(self, iter_, **kwargs):
        print("")
        print("******** iter: %d, iter_time: %.2fs, total_time: %.2fs" %
              (iter_, time.time() - self._check_time, time.time() - self._start_time))
        for key, value in kwargs.items():
            print("%s" % key + "".join([", %s: %.4f" % (k, v) for k, v in value.items()]))
        self._check_time = time.time()

    def save_model(self, save_path):
        self.trainer.save_model(save_path)

    def load_model(self, load_path):
        self.trainer.load_model(load_path)


def _jsonable(v):
    try:
        json.dumps(v)
        return True
    except TypeError:
        return False
