"""Diagnostic: raw gradients of ONE recurrent update (first epoch of a golden's first iteration) on the GPU against the
float64 oracle, per tensor."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from mappo_util import make_params, net_shapes  # noqa: E402
from oracle import mappo_oracle as mo  # noqa: E402
from test_mappo_cuda import build, fill_buffer, load  # noqa: E402
from test_rnn_cuda import fill_rnn  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "rnn_chunk_4x20_h256"
    backend = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    g = load(name)
    c = dict(g["cfg"])
    c["ppo_epoch"] = 1
    if c.get("num_mini_batch", 1) > 1:
        c["num_mini_batch"] = 1
    T, E = g["it1_actions"].shape[:2]
    p = "it1_"
    a_shapes, c_shapes = net_shapes(c)
    orc = mo.Trainer(make_params(a_shapes, c["actor_seed"]), make_params(c_shapes, c["critic_seed"]), c)
    rec = []
    real = mo.clip_grads

    def spy(grads, max_norm, do_clip=True):
        rec.append({k: v.copy() for k, v in grads.items()})
        return real(grads, max_norm, do_clip)
    mo.clip_grads = spy
    n_chunks = T * E * c["n_agents"] // (c["data_chunk_length"] if c["use_recurrent_policy"] else T)
    perm = np.arange(n_chunks)[None]
    orc.train(g[p + "obs"], g[p + "actions"], g[p + "logp"], g[p + "value_preds"], g[p + "returns"], float(g[p + "lr"]), 1,
              perms=perm, rnn_states=g[p + "rnn_states"], rnn_states_critic=g[p + "rnn_states_critic"], masks=g[p + "masks"])
    cfg, pol, tr, buf = build(c, E, T, gemm_backend=backend)
    fill_buffer(buf, g, p)
    fill_rnn(buf, g, p)
    buf.returns_te.copy_(torch.from_numpy(np.ascontiguousarray(g[p + "returns"][:, :, 0, 0])).to(buf.device))
    pol.lr_decay(1, c["n_iters"])
    tr.permutation_fn = lambda ep, n: perm[0].astype(np.int64)
    tr.train(buf)
    for tag, net, og in (("actor", pol.actor, rec[0]), ("critic", pol.critic, rec[1])):
        for k in net.layout:
            got = net.view(k, "grads").detach().cpu().numpy().astype(np.float64)
            want = og[k].reshape(got.shape)
            d = np.abs(got - want)
            i = int(np.argmax(d))
            print("%-6s %-32s max|g| %.3e  max|d| %.3e (rel to max %.1e)  at %d: got %.6e want %.6e" %
                  (tag, k, np.abs(want).max(), d.max(), d.max() / max(np.abs(want).max(), 1e-30), i, got.reshape(-1)[i], want.reshape(-1)[i]))
            if k == "base.feature_norm.weight":
                print("       element 41: got %.6e want %.6e" % (got[41], want[41]))


if __name__ == "__main__":
    main()
