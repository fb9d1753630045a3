"""Harness that runs the UNMODIFIED reference (mounted read-only at /root/reference) in this container.

Test scaffolding, not product code.  Used only by `make_golden.py` (golden-vector generation) and by
the optional `-m "not gpu"` tests that are skipped when /root/reference is absent (it does not exist
on the GPU box).  Nothing here is imported by the product package, `bench.py` or `smoke()`.

Recipe: put `tests/oracle_shim` (stubs for gym / imp / omegaconf /
imageio / wandb) first on sys.path, then `/root/reference/uav_dcc_control`, and import the reference's
own modules by their top-level names (`envs.…`, `algos.…`).

The shipped reference only works at 4 UAV / 20 PoI with the connectivity force off
(envs/mpe/multiagent/scenarios/coverage.py:34,40-41).  `GenScenario` overrides ONLY `make_world`
to lift those literals (Appendix F); every arithmetic method (`CoverageWorld.step`, `Scenario.reward`,
`observation`, `done`, the wrappers, the learner) is the reference's own code.
"""
import os
import sys

import numpy as np

REF_ROOT = os.environ.get("DCC_REFERENCE_ROOT", "/root/reference/uav_dcc_control")
SHIM = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle_shim")


def reference_available():
    return os.path.isdir(REF_ROOT)


_loaded = {}


def load_reference():
    """Import the reference modules (idempotent). Returns a dict of the modules/classes needed."""
    if _loaded:
        return _loaded
    if not reference_available():
        raise RuntimeError("reference not mounted at %s" % REF_ROOT)
    sys.dont_write_bytecode = True  # /root/reference is read-only
    for p in (REF_ROOT, SHIM):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    cwd = os.getcwd()
    os.chdir(REF_ROOT)  # reference uses cwd-relative config paths
    try:
        import envs.mpe.multiagent.scenarios as scenarios
        from envs.mpe.multiagent.CoverageWorld import CoverageWorld
        from envs.mpe.multiagent.core import Agent, Landmark
        from envs.mpe.multiagent.environment import MultiAgentEnv
        from envs.mpe.uav_dcc import DCEnv
        import envs.wrappers as wrappers
    finally:
        os.chdir(cwd)
    mod = scenarios.load("coverage.py")

    class GenScenario(mod.Scenario):
        """Generalised N/M/force: only make_world differs from coverage.py:33-62."""

        reference_compat = False

        def make_world(self):
            if self.reference_compat:
                world = CoverageWorld()  # shipped: scenario comm args never reach the world
            else:
                world = CoverageWorld(self.comm_r_scale, self.comm_force_scale)
            world.dist_mat = np.zeros([self.num_agents, self.num_agents])
            world.collaborative = True
            world.agents = [Agent() for _ in range(self.num_agents)]
            world.landmarks = [Landmark() for _ in range(self.num_pois)]
            for i, agent in enumerate(world.agents):
                agent.name = "agent_%d" % i
                agent.collide = False
                agent.silent = True
                agent.size = self.size
                agent.r_cover = self.r_cover
                agent.r_comm = self.r_comm
                agent.max_speed = 0.5
            for i, landmark in enumerate(world.landmarks):
                landmark.name = "poi_%d" % i
                landmark.collide = False
                landmark.movable = False
                landmark.size = self.size
                landmark.m_energy = self.m_energy
            self.reset_world(world)
            return world

    _loaded.update(dict(scenarios=scenarios, CoverageWorld=CoverageWorld, Agent=Agent, Landmark=Landmark,
                        MultiAgentEnv=MultiAgentEnv, DCEnv=DCEnv, wrappers=wrappers, scenario_mod=mod,
                        GenScenario=GenScenario))
    return _loaded


class RefEnv:
    """One reference env instance + the DummyVecEnv auto-reset rule (envs/wrappers.py:222-235)."""

    def __init__(self, n_agents, n_pois, r_cover=0.2, r_comm=0.4, comm_r_scale=0.95, comm_force_scale=0.0,
                 reference_compat=False, pos_pois=None):
        ref = load_reference()
        sc = ref["GenScenario"](n_agents, n_pois, r_cover, r_comm, comm_r_scale, comm_force_scale)
        sc.reference_compat = reference_compat
        if pos_pois is not None:
            sc.pos_pois = np.asarray(pos_pois, dtype=np.float64)
        self.scenario = sc
        self.world = sc.make_world()
        self.env = ref["MultiAgentEnv"](world=self.world, reset_callback=sc.reset_world,
                                        reward_callback=sc.reward, observation_callback=sc.observation,
                                        done_callback=sc.done)
        self.n_agents, self.n_pois = n_agents, n_pois

    def reset(self):
        return np.array(self.env.reset())

    def state(self):
        w = self.world
        pv = np.array([[a.state.p_pos[0], a.state.p_pos[1], a.state.p_vel[0], a.state.p_vel[1]] for a in w.agents])
        en = np.array([lm.energy for lm in w.landmarks])
        return pv, en

    def set_state(self, pos_vel, energy):
        for a, s in zip(self.world.agents, pos_vel):
            a.state.p_pos = np.array(s[0:2], dtype=np.float64)
            a.state.p_vel = np.array(s[2:4], dtype=np.float64)
        for lm, e in zip(self.world.landmarks, energy):
            lm.energy = float(e)
            lm.done = bool(e >= lm.m_energy)
            lm.just = False

    def step(self, actions_f32):
        """actions (N,2) float32 (copied: the reference scales the caller's array in place).
        Returns a dict holding pre-reset obs, post-auto-reset obs, reward, done, info and world flags."""
        a = np.array(actions_f32, dtype=np.float32, copy=True)
        obs_n, rew_n, done_n, info = self.env.step([a[i] for i in range(self.n_agents)])
        w = self.world
        out = dict(obs_pre=np.array(obs_n), reward=np.array(rew_n, dtype=np.float64),
                   done=np.array(done_n, dtype=bool), coverage_rate=float(w.coverage_rate),
                   connect=bool(w.connect), connect_=bool(w.connect_),
                   adj=np.array(w.adj_mat, dtype=np.uint8), adj_=np.array(w.adj_mat_, dtype=np.uint8))
        out["pos_vel_pre"], out["energy_pre"] = self.state()
        if np.all(done_n):
            out["obs"] = np.array(self.env.reset())
        else:
            out["obs"] = out["obs_pre"]
        out["pos_vel"], out["energy"] = self.state()
        return out
