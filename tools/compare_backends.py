"""Diagnostic: one epoch of gradients with the SIMT fp32 and the tcgen05 3xTF32 GEMM backends on the same synthetic
rollout; prints per-parameter |grad| scale and the max abs difference."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from test_mappo_cuda import build
from dcc_b200 import _lib

N, M, Hd, E, T = 4, 6, 256, 24, 20
D = 4 + 2 * (N - 1) + 5 * M
c = dict(n_agents=N, n_pois=M, hidden=Hd, obs_dim=D, ppo_epoch=1, seed=5, n_iters=10, actor_seed=11, critic_seed=12)
rng = np.random.default_rng(7)
obs = rng.normal(0, 1.5, (T + 1, E, N, D)).astype(np.float32)
act = rng.normal(0, 1.2, (T, E, N, 2)).astype(np.float32)
out = {}
for be in (1, 2):
    rng = np.random.default_rng(8)
    cfg, pol, tr, buf = build(c, E, T, gemm_backend=be)
    dev = buf.device
    buf.obs.copy_(torch.from_numpy(obs).to(dev)); buf.actions.copy_(torch.from_numpy(act).to(dev))
    tr.value_normalizer.state[:3] = torch.tensor([0.3, 4.0, 0.02], device=dev)
    _, logp, _ = pol.evaluate_actions(None, buf.obs[:-1], None, None, buf.actions)
    if be == 1:
        lp_ref = logp.cpu().numpy().reshape(T, E, N).copy()
    lp_old = lp_ref + rng.normal(0, 0.25, (T, E, N)).astype(np.float32)
    vals = rng.normal(0, 1.0, (T + 1, E)).astype(np.float32)
    rew = rng.normal(0, 30.0, (T, E)).astype(np.float32)
    rew[rng.random((T, E)) < 0.05] += 400.0
    rew[rng.random((T, E)) < 0.05] -= 400.0
    masks = (rng.random((T + 1, E)) > 0.1).astype(np.float32)
    for dst, a in ((buf.action_log_probs_ten, lp_old), (buf.values_te, vals), (buf.rewards_te, rew), (buf.masks_te, masks)):
        dst.copy_(torch.from_numpy(a).to(dev))
    buf.compute_returns(None, tr.value_normalizer, policy=pol)
    nep = int(os.environ.get("EPOCHS", "1"))
    tr.ppo_epoch = nep
    tr._epoch_stats = torch.zeros((nep, 4), dtype=torch.float64, device=dev)
    tr._gnorm_sq = torch.zeros((nep, 2), dtype=torch.float64, device=dev)
    lr = float(os.environ.get("LR", "0"))
    pol.lr_actor_now = pol.lr_critic_now = lr
    print(be, tr.train(buf))
    torch.cuda.synchronize()
    which = "grads" if lr == 0 else "params"
    out[be] = {("actor", k): pol.actor.view(k, which).cpu().numpy().copy() for k in pol.actor.layout}
    out[be].update({("critic", k): pol.critic.view(k, which).cpu().numpy().copy() for k in pol.critic.layout})
    v = pol.get_values(buf.obs[:-1].reshape(T * E, N * D)).cpu().numpy()
    out[be]["v"] = v
    out[be]["logp"] = logp.cpu().numpy()
for k in out[1]:
    a, b = out[1][k], out[2][k]
    print("%-50s scale %.3e  maxdiff %.3e  rel %.2e" % (str(k), np.abs(a).max(), np.abs(a - b).max(), np.abs(a - b).max() / max(np.abs(a).max(), 1e-30)))
