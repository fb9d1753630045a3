"""Multi-GPU parity check of the sharded MAPPO update, run under torchrun on N GPUs:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dist_check.py
Every rank builds the same synthetic rollout, trains on ITS contiguous env shard with one NCCL all-reduce of the flat
gradient per epoch, and rank 0 also trains a single-GPU replica on the whole batch.  The two must agree up to fp32
summation order (all-reduce tree + per-rank chunk boundaries)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from dcc_b200.parallel import Comm, init_from_env, shard_envs  # noqa: E402
from test_mappo_cuda import build  # noqa: E402


class _Solo:
    world, rank, calls = 1, 0, 0

    def all_reduce_sum_(self, t):
        return t


def fill(buf, obs, act, lp, vals, rew, masks, lo, hi):
    dev = buf.device
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)   # noqa: E731
    buf.obs.copy_(t(obs[:, lo:hi])); buf.actions.copy_(t(act[:, lo:hi])); buf.action_log_probs_ten.copy_(t(lp[:, lo:hi]))
    buf.values_te.copy_(t(vals[:, lo:hi])); buf.rewards_te.copy_(t(rew[:, lo:hi])); buf.masks_te.copy_(t(masks[:, lo:hi]))


def main():
    comm = init_from_env()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    ok = True
    for nmb in (1, 2):      # whole-rollout minibatch, and num_mini_batch = 2 (per-rank permutations, shared_buffer.py:219-279)
        ok = run(comm, local, nmb) and ok
    if comm.rank == 0 and not ok:
        sys.exit(1)


def run(comm, local, nmb):
    N, M, Hd, E, T, EPOCHS = 8, 64, 256, 256, 12, 3
    D = 4 + 2 * (N - 1) + 5 * M
    c = dict(n_agents=N, n_pois=M, hidden=Hd, obs_dim=D, ppo_epoch=EPOCHS, seed=5, n_iters=10, actor_seed=11, critic_seed=12,
             num_mini_batch=nmb)
    # every rank's local permutations (all ranks can compute all of them) and the equivalent global permutation: global
    # minibatch i = union over ranks of that rank's minibatch i, local agent row (t, e, n) -> global ((t*E + lo + e)*N + n)
    W = comm.world
    shards = [shard_envs(E, W, r) for r in range(W)]
    perms_local = [[np.random.default_rng(1000 * r + ep).permutation(T * (hi_ - lo_) * N) for ep in range(EPOCHS)]
                   for r, (lo_, hi_) in enumerate(shards)]

    def to_global(r, idx):
        lo_, hi_ = shards[r]
        El = hi_ - lo_
        t, rem = np.divmod(idx, El * N)
        e, n = np.divmod(rem, N)
        return (t * E + lo_ + e) * N + n
    perms_global = []
    for ep in range(EPOCHS):
        parts = []
        for i in range(nmb):
            for r, (lo_, hi_) in enumerate(shards):
                mbs = T * (hi_ - lo_) * N // nmb
                parts.append(to_global(r, perms_local[r][ep][i * mbs:(i + 1) * mbs]))
        perms_global.append(np.concatenate(parts))
    rng = np.random.default_rng(3)
    obs = rng.normal(0, 1.2, (T + 1, E, N, D)).astype(np.float32)
    act = rng.normal(0, 1.0, (T, E, N, 2)).astype(np.float32)
    vals = rng.normal(0, 1.0, (T + 1, E)).astype(np.float32)
    rew = rng.normal(0, 20.0, (T, E)).astype(np.float32)
    masks = (rng.random((T + 1, E)) > 0.05).astype(np.float32)
    lo, hi = shard_envs(E, comm.world, comm.rank)
    out = {}
    for tag, e_lo, e_hi, cm in (("sharded", lo, hi, comm),) + ((("solo", 0, E, _Solo()),) if comm.rank == 0 else ()):
        cfg, pol, tr, buf = build(c, e_hi - e_lo, T, device=local)
        tr.comm = cm
        buf.n_envs_global = E
        tr.permutation_fn = (lambda ep, n: perms_local[comm.rank][ep]) if tag == "sharded" else (lambda ep, n: perms_global[ep])
        tr.value_normalizer.state[:3] = torch.tensor([0.3, 4.0, 0.02], device=buf.device)
        fill(buf, obs, act, np.zeros((T, E, N), np.float32), vals, rew, masks, e_lo, e_hi)
        _, logp, _ = pol.evaluate_actions(None, buf.obs[:-1], None, None, buf.actions)
        noise = np.random.default_rng(9).normal(0, 0.2, (T, E, N)).astype(np.float32)[:, e_lo:e_hi]
        buf.action_log_probs_ten.copy_(logp.view(T, e_hi - e_lo, N) + torch.from_numpy(noise).to(buf.device))
        buf.compute_returns(None, tr.value_normalizer, policy=pol)
        pol.lr_decay(2, 10)
        info = tr.train(buf)
        torch.cuda.synchronize()
        out[tag] = (info, pol.actor.params.cpu().numpy(), pol.critic.params.cpu().numpy(),
                    tr.value_normalizer.state.cpu().numpy()[:3], cm.calls)
    comm.barrier()
    if comm.rank == 0:
        a, b = out["sharded"], out["solo"]
        res = {"world": comm.world, "epochs": EPOCHS, "num_mini_batch": nmb, "allreduce_calls": a[4], "backend": pol.gemm_backend()}
        for k in a[0]:
            res["info_" + k] = [a[0][k], b[0][k]]
        lr = pol.lr_actor_now
        for name, x, y in (("actor", a[1], b[1]), ("critic", a[2], b[2])):
            d = np.abs(x - y)
            res[name + "_max_abs_diff"] = float(d.max())
            res[name + "_frac_off"] = float((d > 1e-5 + 2e-5 * np.abs(y)).mean())
        res["valuenorm_equal"] = bool(np.allclose(a[3], b[3], rtol=1e-6))
        ok = all(abs(u - v) <= 5e-5 * max(1.0, abs(v)) for u, v in (res["info_" + k] for k in a[0]))
        ok = ok and res["actor_frac_off"] < 2e-3 and res["critic_frac_off"] < 2e-3 and res["valuenorm_equal"]
        ok = ok and res["actor_max_abs_diff"] <= 6 * lr and res["critic_max_abs_diff"] <= 6 * lr
        res["ok"] = bool(ok)
        print(json.dumps(res), flush=True)
        return bool(ok)
    return True


if __name__ == "__main__":
    main()
