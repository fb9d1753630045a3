#!/usr/bin/env python
"""bench.py — agent-steps/s of the env hot path at 8 UAV / 64 PoI / 65 536 envs per GPU (BASELINE.json), with the full
MAPPO loop (configs[3] / [4]) measured in the same run.

    python bench.py [--gpus N] [--steps K] [--warmup W]              # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]   # the reference on the host cores
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...  # one rank per GPU, envs sharded (weak scaling)

Workloads (`--workload`):
  env    (default, the BASELINE.json metric: configs[1]) one "step" = one `env.step` of all E env instances of a rank
         (one kernel launch).  The line also carries `full_loop`: ONE timed iteration (after one warm-up iteration) of the
         full MAPPO loop at 65 536 envs per GPU — 150-step rollout, GAE, 15-epoch PPO update, and under torchrun the NCCL
         all-reduce of the flat gradient once per epoch (configs[3] on 1 GPU, configs[4] sharded) — so that the driver's
         BENCH / SCALE runs exercise the learner and the collective too (`--no-full-loop` skips it);
  env16  BASELINE configs[2]: 16 UAV / 256 PoI / 32 768 envs, connectivity force on, env step only;
  mappo  the full loop as the line's own metric (one "step" = one iteration; default --steps 2 --warmup 3).
Prints ONE JSON line:
  value        whole-job agent-steps/s, inputs (actions) resident in HBM, device-timed with CUDA events
  e2e          same metric through the host-buffer C entry point `dcc_env_step_host` (numpy in / numpy out):
               pinned H2D of the actions and D2H of obs/reward/done/coverage inside the timed region
  roofline     algorithmic bytes per launch / average launch time vs the measured HBM copy bandwidth
  cpu_baseline the CPU oracle port (oracle/dcc_env_oracle.c, kind "port") timed on this box's host cores, bounded
               sample; `cpu_baseline.python_reference` = the UNMODIFIED Python reference itself timed on the same cores
               in the same run (baseline/run_reference.py: SubprocVecEnv x cpu_count env-only at 8/64, and the shipped
               4/20 full loop), from /root/reference or the verbatim copy under baseline/_ref
  full_loop    see above
`--impl reference` times the unmodified Python reference (kind "reference"; the C port's number rides along).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_AGENTS, N_POIS, ENVS_PER_GPU = 8, 64, 65536
METRIC = "agent-steps/sec at 8 UAV / 64 PoI (env step, 65536 envs per GPU)"
UNIT = "agent-steps/s"


def obs_dim(n, m):
    return 4 + 2 * (n - 1) + 5 * m


def alg_bytes_per_agent_step(n, m):
    """obs write 4D + action read 8 + reward 4 + done 1 + pos/vel f64 r+w 64
    + per env (PoI energy u8 r+w 2M, coverage_rate 4, connect 1) / N."""
    return 4 * obs_dim(n, m) + 8 + 4 + 1 + 64 + (2 * m + 5) / n


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks / throttle reasons DURING the timed region from a background thread (NVML, ~2 ms period;
    the timed region is tens of ms, too short for `nvidia-smi -lms`).  Same fields as the B200_PROFILING.md recipe."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.err = index, [], set(), None
        self._stop, self._thr, self.max_mhz = False, None, None

    def _run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # honour CUDA_VISIBLE_DEVICES when it lists plain indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis:
                try:
                    idx = int(vis.split(",")[self.index])
                except (ValueError, IndexError):
                    pass
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                     "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                     "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
            while not self._stop:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
                time.sleep(0.002)
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def start(self):
        import threading
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()
        time.sleep(0.05)  # NVML init outside the timed region

    def mark(self):
        """Drop samples taken before the timed region starts."""
        self.samples.clear()
        self.reasons.clear()

    def stop(self):
        self._stop = True
        if self._thr is not None:
            self._thr.join(timeout=2)
        out = {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
               "samples": len(self.samples), "reasons": sorted(self.reasons), "source": "NVML, sampled during the timed region"}
        if self.err:
            out["error"] = self.err
        return out


def time_cpu_port(n_envs, steps, warmup, threads, seed=0, N=N_AGENTS, M=N_POIS, env_kw=None):
    """Times the CPU oracle port (test infrastructure; here only as the measured CPU baseline)."""
    from oracle.env_oracle import OracleEnv
    from dcc_b200.envs.cuda_vec_env import synthetic_pois
    env_kw = env_kw or {}
    if env_kw.get("reference_compat", True):       # shipped semantics: the world ignores the scenario's comm arguments
        crs, force = 0.9, 0.0
    else:
        crs, force = env_kw.get("comm_r_scale", 0.95), 100.0 * env_kw.get("comm_force_scale", 0.0)
    env = OracleEnv(n_envs, N, M, synthetic_pois(M), comm_r_scale=crs, contact_force=force, n_threads=threads)
    rng = np.random.default_rng(seed)
    acts = [rng.standard_normal((n_envs, N, 2)).astype(np.float32) for _ in range(4)]
    env.reset()
    for t in range(warmup):
        env.step(acts[t % 4], want_aux=False)
    t0 = time.perf_counter()
    for t in range(steps):
        env.step(acts[t % 4], want_aux=False)
    dt = time.perf_counter() - t0
    return n_envs * N * steps / dt, dt


def time_python_reference(what, timeout_s=240, **kw):
    """Runs baseline/run_reference.py (the UNMODIFIED Python reference behind the stub shim) in a subprocess on the
    host cores and returns its JSON record, or {"unavailable": why}."""
    cmd = [sys.executable, os.path.join(ROOT, "baseline", "run_reference.py"), what]
    for k, v in kw.items():
        cmd += ["--" + k, str(v)]
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):   # not a torchrun worker
        env.pop(k, None)
    env["CUDA_VISIBLE_DEVICES"] = ""        # the reference's CPU path (ptu.set_gpu_mode(False))
    try:
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s, env=env, cwd=ROOT)
        lines = [l for l in res.stdout.strip().splitlines() if l.startswith("{")]
        if res.returncode != 0 or not lines:
            return {"unavailable": "run_reference.py %s failed (rc %d): %s" % (what, res.returncode, res.stderr.strip()[-300:])}
        return json.loads(lines[-1])
    except subprocess.TimeoutExpired:
        return {"unavailable": "run_reference.py %s exceeded %d s" % (what, timeout_s)}
    except Exception as e:  # pragma: no cover
        return {"unavailable": repr(e)}


def python_reference_record(env_steps=150, loop_iters=2, n=N_AGENTS, m=N_POIS):
    """cpu_baseline.python_reference: SURVEY.md §8d (i) env-only at the benchmarked shape with SubprocVecEnv x cpu_count,
    (ii) the shipped 4 UAV / 20 PoI full loop (16 SubprocVecEnv workers, T = 150, 15 epochs)."""
    cores = min(os.cpu_count() or 1, 64)        # one OS process per env instance (SubprocVecEnv): bounded on very wide hosts
    out = {"cores": cores, "kind": "reference",
           "how": "unmodified reference sources (baseline/run_reference.py; stub shim for gym/imp/omegaconf/imageio/wandb; "
                  "make_world's 4/20 literals lifted for 8/64)"}
    out["env_only"] = time_python_reference("env", timeout_s=150, n=n, m=m, procs=cores, steps=env_steps, warmup=10)
    if loop_iters > 0:
        out["full_loop_4x20_shipped"] = time_python_reference("loop", iters=loop_iters)
    return out


def run_reference(args):
    """Reference arm: the UNMODIFIED Python reference, env-step-only at 8 UAV / 64 PoI through its own
    SubprocVecEnv (one OS process per env, as many envs as host cores) — a bounded sample of the 65 536-env workload,
    throughput-normalised.  The C port of the same algorithm (all host threads) rides along in cpu_baseline.port.
    If the reference sources are not on the box the line falls back to the port (kind "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.env_oracle import max_threads
    threads = max_threads()
    cores = min(os.cpu_count() or 1, 64)
    n_envs = 8192
    port_value, port_dt = time_cpu_port(n_envs, max(args.steps, 30), min(args.warmup, 3), threads)
    port = {"value": port_value, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d envs x %d steps of the 8 UAV / 64 PoI env step, float64 C port of the reference algorithm "
                      "(oracle/dcc_env_oracle.c)" % (n_envs, max(args.steps, 30))}
    steps = max(10, min(args.steps, 150))
    rec = time_python_reference("env", n=N_AGENTS, m=N_POIS, procs=cores, steps=steps, warmup=max(3, min(args.warmup, 10)))
    if "agent_steps_per_s" in rec:
        value, dt, kind = rec["agent_steps_per_s"], rec["seconds"], "reference"
        sample = ("%d envs (SubprocVecEnv, one process per env) x %d steps of the 8 UAV / 64 PoI env step: the unmodified "
                  "Python reference from %s" % (rec["procs"], steps, rec["reference_root"]))
        envs, used = rec["procs"], cores
    else:
        value, dt, kind, sample, envs, used, steps = port_value, port_dt, "port", port["sample"], n_envs, threads, max(args.steps, 30)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "env-step-only, 8 UAV / 64 PoI, %d envs per CPU step (bounded sample of the 65536-env "
                               "workload)" % envs, "n_agents": N_AGENTS, "n_pois": N_POIS, "envs": envs},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": kind, "sample": sample, "port": port,
                         "python_reference": rec},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


WORKLOADS = {
    # name: (N, M, envs per GPU, env kwargs, description, kernel)
    "env": (8, 64, 65536, dict(reference_compat=True),
            "env-step-only (BASELINE configs[1]): 8 UAV / 64 PoI, 65536 envs per GPU, N(0,1) float32 actions, auto-reset on, "
            "shipped semantics (reference_compat)", "dcc_env_spec_kernel<8,64,true>"),
    "env16": (16, 256, 32768, dict(reference_compat=False, comm_r_scale=0.95, comm_force_scale=1.0),
              "env-step-only (BASELINE configs[2]): 16 UAV / 256 PoI, 32768 envs per GPU, connectivity constraint active "
              "(comm_r_scale 0.95, contact force 100), N(0,1) float32 actions, auto-reset on", "dcc_env_spec_kernel<16,256,true>"),
}


def run_cuda(args):
    import torch
    import torch.distributed as dist
    from dcc_b200.envs import CudaVecEnv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    N, M, E, env_kw, workload_desc, kernel_name = WORKLOADS[args.workload]
    D = obs_dim(N, M)
    env = CudaVecEnv(E, N, M, device=local_rank, pos_pois="synthetic", **env_kw)   # synthetic PoI layout (uniform, seed 0)
    if args.per_env_layouts:
        env.set_poi_layouts(np.random.default_rng(rank).uniform(-1.0, 1.0, (E, M, 2)))
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    acts = [torch.randn((E, N, 2), generator=gen, device=dev, dtype=torch.float32) for _ in range(8)]
    env.reset()
    for t in range(max(args.warmup, 3)):
        env.step(acts[t % 8])
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: K steps, CUDA events on the launching (current) stream ----------------
    sampler = ClockSampler(local_rank)
    launches0 = env.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if rank == 0:
        sampler.start()
    barrier()
    if rank == 0:
        sampler.mark()
    ev0.record()
    for t in range(args.steps):
        env.step(acts[t % 8])
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    launches = env.launch_count() - launches0
    if world > 1:
        tt = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    value = world * E * N * args.steps / (ms * 1e-3)

    # ---- end to end through the host-buffer C entry point -----------------------------------------------
    e2e_steps = max(3, min(args.steps, 20))
    hb = env._host_buffers()
    host_acts = [a.cpu().numpy() for a in acts[:4]]
    for t in range(2):
        env.step_host(host_acts[t % 4])
    barrier()
    t0 = time.perf_counter()
    for t in range(e2e_steps):
        np.copyto(hb["actions"], host_acts[t % 4])
        env.step_host(hb["actions"])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e_value = world * E * N * e2e_steps / e2e_s
    h2d = E * N * 2 * 4
    d2h = E * N * D * 4 + E * N * 4 + E * N + E * 4
    env.close()
    del env, acts
    torch.cuda.empty_cache()

    # ---- the full MAPPO loop in the same run (configs[3] / [4]): every rank takes part --------------------
    full = None
    if args.workload == "env" and not args.no_full_loop:
        full = measure_full_loop(args.envs or ENVS_PER_GPU, 1, 1, world, rank, local_rank, dev)

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        b_alg = alg_bytes_per_agent_step(N, M) + (16.0 * M / N if args.per_env_layouts else 0.0)
        launch_s = ms * 1e-3 / args.steps
        achieved = E * N * b_alg / launch_s / 1e9
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "env_step_traffic.json" if args.workload == "env" else "env16_step_traffic.json")
        if os.path.exists(tp) and not args.per_env_layouts:     # the ncu capture is of the shared-layout launch
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
                traffic_src = ("ncu --set full capture of this kernel at this shape, committed as profiles/%s — a constant "
                               "read from that file, NOT measured in this run" % os.path.basename(tp))
            except Exception:
                traffic = None
        cpu = None
        if not args.no_cpu_baseline:
            from oracle.env_oracle import max_threads
            threads = max_threads()
            n_envs_cpu, steps_cpu = (8192, 30) if args.workload == "env" else (1024, 10)
            v, dt = time_cpu_port(n_envs_cpu, steps_cpu, 2, threads, N=N, M=M, env_kw=env_kw)
            # grow the sample to >= ~10 s of CPU work
            if dt < 10.0:
                steps_cpu = int(min(2000, steps_cpu * 10.0 / max(dt, 1e-3)))
                v, dt = time_cpu_port(n_envs_cpu, steps_cpu, 0, threads, N=N, M=M, env_kw=env_kw)
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": "%d envs x %d steps (%.1f s) of the same %d/%d env step, float64 C port of the reference "
                             "algorithm (oracle/dcc_env_oracle.c), %d POSIX threads" % (n_envs_cpu, steps_cpu, dt, N, M, threads)}
            if world == 1 and not args.no_python_reference:
                cpu["python_reference"] = python_reference_record(env_steps=150 if args.workload == "env" else 40,
                                                                  loop_iters=2 if args.workload == "env" else 0, n=N, m=M)
        metric = METRIC if args.workload == "env" else "agent-steps/sec at 16 UAV / 256 PoI (env step, 32768 envs per GPU, connectivity force on)"
        line = {
            "metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_desc,
                       "n_agents": N, "n_pois": M, "envs_per_gpu": E, "obs_dim": D, "poi_layout": ("per-env uniform(-1,1) (dcc_env_set_poi_layouts)" if args.per_env_layouts
                                      else "uniform(-1,1), seed 0"),
                       "l2": "no flush needed: %.0f MB written per step > 126 MB L2" % (E * N * D * 4 / 1e6)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "alg_bytes_per_agent_step": b_alg, "alg_bytes_per_launch": E * N * b_alg,
                         "frac_obs_bytes_only": E * N * 4.0 * D / launch_s / 1e9 / peak,
                         "frac_dram_traffic": (traffic / launch_s / 1e9 / peak) if traffic else None,
                         "kernel": kernel_name, "avg_launch_us": launch_s * 1e6},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": "CudaVecEnv.step_host -> dcc_env_step_host (pinned host buffers)"},
            "gpu_launches": launches + (full["gpu_launches"] if full else 0),
            "clocks": clocks,
        }
        if full is not None:
            line["full_loop"] = full
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()



# ---- MAPPO full-loop workload (BASELINE configs[3] single GPU, configs[4] sharded) -----------------------------------
MAPPO_METRIC = "agent-steps/sec at 8 UAV / 64 PoI (full MAPPO rollout+update loop, 65536 envs per GPU)"
T_ROLLOUT, PPO_EPOCH, HIDDEN = 150, 15, 256


def mappo_flops_per_env_step_row(n, m, hidden=HIDDEN):
    """Algorithmic fp32 FLOPs of ONE PPO epoch per env-step row (N actor rows + 1 critic row), as this build computes
    it (critic evaluated once per env; no dX GEMM for layer 1): forward 2 GEMMs + backward dX2, dW2, dW1 per net."""
    d = obs_dim(n, m)
    per_net = lambda k: 2 * hidden * (k + hidden) + 2 * hidden * (hidden + hidden + k)   # noqa: E731
    return n * per_net(d) + per_net(n * d)


def time_cpu_port_mappo(n_envs, iters, seed=0):
    """CPU arm of the full loop: the C env oracle + the float64 NumPy MAPPO oracle (hand-written backward), i.e. the
    reference's algorithm on the host cores.  Bounded sample: n_envs envs, T = 150, 15 epochs."""
    import numpy as np
    from oracle.env_oracle import OracleEnv, max_threads
    from oracle import mappo_oracle as mo
    from dcc_b200.envs.cuda_vec_env import synthetic_pois
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from mappo_util import actor_param_shapes, critic_param_shapes, make_params
    N, M, D = N_AGENTS, N_POIS, obs_dim(N_AGENTS, N_POIS)
    hp = dict(clip_param=0.2, entropy_coef=0.01, value_loss_coef=1.0, max_grad_norm=10.0, huber_delta=10.0, opti_eps=1e-5)
    tr = mo.Trainer(make_params(actor_param_shapes(D, HIDDEN), 1), make_params(critic_param_shapes(N * D, HIDDEN), 2), hp)
    env = OracleEnv(n_envs, N, M, synthetic_pois(M), comm_r_scale=0.9, contact_force=0.0, n_threads=max_threads())
    rng = np.random.default_rng(seed)
    T = T_ROLLOUT
    t0 = time.perf_counter()
    for _ in range(iters):
        obs = np.zeros((T + 1, n_envs, N, D), np.float32)
        act = np.zeros((T, n_envs, N, 2), np.float32)
        logp = np.zeros((T, n_envs, N, 1), np.float32)
        vals = np.zeros((T + 1, n_envs, N, 1), np.float32)
        rew = np.zeros((T, n_envs, N, 1), np.float32)
        masks = np.ones((T + 1, n_envs, N, 1), np.float32)
        obs[0] = env.reset()
        logstd = tr.actor.p["act.action_out.logstd._bias"].reshape(1, -1)
        for t in range(T):
            mean = tr.actor.forward(obs[t].reshape(n_envs * N, D))
            a = mean + np.exp(logstd) * rng.standard_normal(mean.shape)
            lp, _ = mo.gaussian_logp_entropy(mean, logstd, a)
            v = tr.critic.forward(obs[t].reshape(n_envs, N * D))
            act[t], logp[t] = a.reshape(n_envs, N, 2), lp.reshape(n_envs, N, 1)
            vals[t] = np.repeat(v.reshape(n_envs, 1, 1), N, axis=1)
            o = env.step(act[t], want_aux=False)
            obs[t + 1] = o["obs"]
            rew[t] = o["reward"].reshape(n_envs, 1, 1)
            masks[t + 1] = 1.0 - o["done"].reshape(n_envs, 1, 1)
        vals[T] = np.repeat(tr.critic.forward(obs[T].reshape(n_envs, N * D)).reshape(n_envs, 1, 1), N, axis=1)
        ret = mo.gae_returns(rew, vals, masks, tr.vn)
        tr.train(obs, act, logp, vals, ret, 5e-4, PPO_EPOCH)
    dt = time.perf_counter() - t0
    return n_envs * N * T * iters / dt, dt


def run_reference_mappo(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    n_envs, iters = 4, max(1, min(args.steps, 2))
    value, dt = time_cpu_port_mappo(n_envs, iters)
    threads = os.cpu_count() or 1
    sample = ("%d envs x %d iterations (T=150, 15 epochs, %.1f s) of the 8 UAV / 64 PoI MAPPO loop: C env oracle + float64 "
              "NumPy MAPPO oracle (BLAS threads: all host cores)" % (n_envs, iters, dt))
    line = {"impl": "reference", "metric": MAPPO_METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": iters,
            "warmup": 0, "ms_per_step": 1e3 * dt / iters, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "full MAPPO loop, 8 UAV / 64 PoI, bounded sample of %d envs" % n_envs, "n_agents": N_AGENTS,
                       "n_pois": N_POIS, "envs": n_envs, "T": T_ROLLOUT, "ppo_epoch": PPO_EPOCH},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def mappo_flops_compact(n, m, hidden=HIDDEN):
    """Algorithmic fp32 FLOPs of ONE PPO epoch per env-step row with the compact-state first layer (the path the loop
    runs): K = 2N + 4 + 2M for an actor row, N (2N + 2) + 2M + 2 for the critic row (csrc/dcc_compact.cuh)."""
    ka, kc = 2 * n + 4 + 2 * m, n * (2 * n + 2) + 2 * m + 2
    per_net = lambda k: 2 * hidden * (k + hidden) + 2 * hidden * (hidden + hidden + k)   # noqa: E731
    return n * per_net(ka) + per_net(kc)


def measure_full_loop(E, steps, warmup, world, rank, local_rank, dev):
    """`steps` timed iterations (after `warmup`) of the full MAPPO loop at E envs per GPU: T = 150 rollout (policy forward +
    env step + insert), GAE, 15-epoch PPO update; under torchrun the env axis is sharded and the flat actor+critic
    gradient is all-reduced over NCCL once per epoch.  Returns the record (identical on every rank; rank 0 prints it)."""
    import torch
    import torch.distributed as dist
    from dcc_b200.learner import Learner
    from dcc_b200.parallel import Comm
    from dcc_b200.utils.config import load_config

    cfg = load_config(None, num_agents=N_AGENTS, num_pois=N_POIS, n_rollout_threads=E * world, max_ep_len=T_ROLLOUT,
                      ppo_epoch=PPO_EPOCH, n_iters=steps + warmup + 1, n_eval_rollout_threads=0, n_render_rollout_threads=0,
                      save_model=False, device=local_rank, poi_layout="synthetic")
    comm = Comm()
    lr = Learner(cfg, comm=comm)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warmup):
        lr.policy.lr_decay(i + 1, cfg.n_iters)
        lr.rollout(lr.rl_buffer, lr.train_envs)
        lr.rl_update()
    sampler = ClockSampler(local_rank)
    l0 = lr.policy.launch_count() + lr.train_envs.launch_count()
    c0, b0 = comm.calls, comm.bytes
    comm.timing = True
    comm.collective_ms()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * steps + 1)]
    if rank == 0:
        sampler.start()
    barrier()
    if rank == 0:
        sampler.mark()
    t0 = time.perf_counter()
    ev[0].record()
    info = None
    for i in range(steps):
        lr.policy.lr_decay(warmup + i + 1, cfg.n_iters)
        ri = lr.rollout(lr.rl_buffer, lr.train_envs)     # returns python floats: the rollout's device->host read
        ev[2 * i + 1].record()
        ti = lr.rl_update()                              # returns python floats: the update's device->host read
        ev[2 * i + 2].record()
        info = (ri, ti)
    barrier()
    wall_s = time.perf_counter() - t0
    ms = ev[0].elapsed_time(ev[-1])
    roll_ms = sum(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(steps))
    upd_ms = sum(ev[2 * i + 1].elapsed_time(ev[2 * i + 2]) for i in range(steps))
    coll_ms = comm.collective_ms()
    comm.timing = False
    clocks = sampler.stop() if rank == 0 else None
    launches = lr.policy.launch_count() + lr.train_envs.launch_count() - l0 + steps * T_ROLLOUT   # + insert kernels
    if world > 1:
        tt = torch.tensor([ms, wall_s, upd_ms, roll_ms, coll_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, wall_s, upd_ms, roll_ms, coll_ms = (float(x) for x in tt.tolist())
    agent_steps = world * E * N_AGENTS * T_ROLLOUT * steps
    compact = bool(lr.compact)
    peak_tf = 1404.7
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak_tf = float(json.load(f).get("bf16_tflops_sustained", peak_tf))
        peak_src = ("measured sustained dense bf16 (MEASURED_PEAKS.json); fp32-level parity needs three split MMAs per product: "
                    "the fp16 hi/lo split kernels top out at 1/3 of it, the 3xTF32 kernels at 1/6")
    except Exception:
        peak_src = "fallback"
    hbm_peak, _ = measured_peak_gbs()
    flop_row = mappo_flops_compact(N_AGENTS, N_POIS) if compact else mappo_flops_per_env_step_row(N_AGENTS, N_POIS)
    upd_flops = flop_row * float(E) * T_ROLLOUT * PPO_EPOCH * steps          # per rank
    achieved = upd_flops / (upd_ms * 1e-3) / 1e12
    chunk_rows = int(lr.policy.lib.dcc_mappo_chunk_rows(lr.policy._h))
    chunks = steps * PPO_EPOCH * (-(-(E * T_ROLLOUT) // chunk_rows))
    traffic, traffic_src, hbm = None, None, None
    tp = os.path.join(ROOT, "profiles", "mappo_chunk_traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            if bool(tj.get("compact")) == compact and int(tj.get("chunk_rows", 0)) > 0:
                # the capture is per 37 888 env steps; every kernel of the update streams its rows once, so the traffic of a chunk
                # scales with its row count
                traffic = tj["dram_bytes_per_chunk"] * chunk_rows / float(tj["chunk_rows"])
                traffic_src = ("sum of dram__bytes_read+write over the kernels of ONE activation chunk (ncu --set full, captured at %d "
                               "env steps per chunk and scaled to this run's %d), committed as profiles/mappo_chunk_traffic.json — a "
                               "constant read from that file, NOT measured in this run" % (int(tj["chunk_rows"]), chunk_rows))
                hbm = {"bound": "hbm", "achieved": traffic * chunks / (upd_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                       "frac": traffic * chunks / (upd_ms * 1e-3) / 1e9 / hbm_peak, "traffic_per_chunk": traffic,
                       "traffic_source": traffic_src}
        except Exception:
            traffic = None
    buf = lr.rl_buffer
    rollout_bytes = (buf.state_pv.numel() * 8 + buf.state_en.numel()) if compact else buf.obs.numel() * 4
    rec = {
        "metric": MAPPO_METRIC, "value": agent_steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms / steps, "rollout_ms": roll_ms / steps, "update_ms": upd_ms / steps,
        "allreduce_calls": (comm.calls - c0) / steps, "allreduce_ms": coll_ms / steps,
        "allreduce_bytes": (comm.bytes - b0) / steps,
        "allreduce_us_per_epoch": 1e3 * coll_ms / steps / PPO_EPOCH,
        "config": {"workload": "full MAPPO loop (BASELINE configs[3]%s): 8 UAV / 64 PoI, %d envs per GPU, T=150 rollout + GAE + "
                               "15-epoch PPO update, shipped hyper-parameters, random-init policy, synthetic PoI layout" % (
                                   "; configs[4] sharded, NCCL gradient all-reduce once per epoch" if world > 1 else "", E),
                   "n_agents": N_AGENTS, "n_pois": N_POIS, "envs_per_gpu": E, "T": T_ROLLOUT, "ppo_epoch": PPO_EPOCH,
                   "hidden": HIDDEN, "gemm_backend": lr.policy.gemm_backend(), "chunk_rows": chunk_rows,
                   "rollout_storage": ("compact state (%.2f GB: pos/vel float64 + PoI energy uint8 per env step; first layer "
                                       "evaluated from it, csrc/dcc_compact.cuh)" % (rollout_bytes / 1e9)) if compact else
                                      ("materialised float32 observations (%.1f GB)" % (rollout_bytes / 1e9)),
                   "l2": "no flush needed: one epoch streams %.1f GB of rollout + activations" % (
                       (traffic * chunks / steps / PPO_EPOCH / 1e9) if traffic else rollout_bytes / 1e9)},
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "alg_flops_per_env_step_row_per_epoch": flop_row,
                     "alg_flops_per_env_step_row_per_epoch_observation_rows": mappo_flops_per_env_step_row(N_AGENTS, N_POIS),
                     "kernel": "tc_gemm_fwd_kernel + tc_gemm_wgrad_kernel (update phase, fp32-equivalent FLOPs of the GEMMs as "
                               "this build evaluates them)", "hbm": hbm},
        "e2e": {"value": agent_steps / wall_s, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 9 * 8,
                "steps": steps, "api": "Learner.rollout + Learner.rl_update (host wall clock incl. the per-iteration device->host "
                                       "reads of rollout_info / train_info; actions are produced on the device by the policy)"},
        "gpu_launches": int(launches), "clocks": clocks,
        "last_iter": {"rollout_info": info[0], "train_info": info[1]},
    }
    lr.train_envs.close()
    del lr
    torch.cuda.empty_cache()
    return rec


def run_cuda_mappo(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rec = measure_full_loop(args.envs or ENVS_PER_GPU, args.steps, max(args.warmup, 1), world, rank, local_rank, dev)
    if rank == 0:
        line = dict(rec)
        line.update({"higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                     "cpu_baseline": None})
        if not args.no_cpu_baseline:
            v, dt = time_cpu_port_mappo(2, 1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                                    "sample": "2 envs x 1 iteration (T=150, 15 epochs, %.1f s): C env oracle + float64 NumPy MAPPO oracle" % dt}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--workload", default="env", choices=["env", "env16", "mappo"])
    ap.add_argument("--envs", type=int, default=0, help="envs per GPU of the full MAPPO loop (default 65536)")
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-python-reference", action="store_true", help="skip timing the Python reference itself")
    ap.add_argument("--no-full-loop", action="store_true", help="env workload: skip the full_loop sub-record")
    ap.add_argument("--per-env-layouts", action="store_true",
                    help="env workload stress variant (SURVEY.md §8d): every env instance has its own uniform(-1,1) PoI layout "
                         "(+16 M / N bytes read per agent-step)")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 2 if args.workload == "mappo" else 150
    if args.warmup is None:
        args.warmup = 1 if args.workload == "mappo" else 10
    if args.workload == "mappo":
        (run_reference_mappo if args.impl == "reference" else run_cuda_mappo)(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
