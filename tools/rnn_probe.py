"""Diagnostic replay of the recurrent goldens on the GPU: per case / backend / chunk the worst forward, train_info and
per-tensor parameter differences (what tests/test_rnn_cuda.py asserts, printed instead of asserted)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from test_mappo_cuda import ALL_CASES, build, fill_buffer, load  # noqa: E402
from test_rnn_cuda import fill_rnn  # noqa: E402


def main():
    for name in [n for n in ALL_CASES if n.startswith("rnn_")]:
        for backend in (1, 0):
            for chunk in (0, 40):
                g = load(name)
                c = g["cfg"]
                N, D, R, Hd = c["n_agents"], c["obs_dim"], c["recurrent_N"], c["hidden"]
                T, E = g["it1_actions"].shape[:2]
                L = c["data_chunk_length"] if c["use_recurrent_policy"] else T
                if chunk and chunk < L:
                    continue
                cfg, pol, tr, buf = build(c, E, T, gemm_backend=backend, chunk_rows=chunk)
                for it in range(1, c["iters"] + 1):
                    p = "it%d_" % it
                    fill_buffer(buf, g, p)
                    fill_rnn(buf, g, p)
                    logp = torch.empty(E * N, device=buf.device)
                    vals = torch.empty(E, device=buf.device)
                    dl = dv = dh = 0.0
                    for t in range(T):
                        ha, hc = pol._act_rnn(buf.obs[t].contiguous(), E, buf.rnn_a[t], buf.rnn_c[t], buf.masks_te[t], 1, False,
                                              buf.actions[t].contiguous(), logp, vals)
                        dl = max(dl, np.abs(logp.cpu().numpy().reshape(E, N, 1) - g[p + "logp"][t]).max())
                        dv = max(dv, np.abs(vals.cpu().numpy().reshape(E, 1, 1) - g[p + "value_preds"][t][:, :1]).max())
                        keep = g[p + "masks"][t + 1].reshape(E * N, 1, 1)
                        dh = max(dh, np.abs(ha.cpu().numpy() * keep - g[p + "rnn_states"][t + 1].reshape(E * N, R, Hd)).max())
                    buf.returns_te.copy_(torch.from_numpy(np.ascontiguousarray(g[p + "returns"][:, :, 0, 0])).to(buf.device))
                    pol.lr_decay(it, c["n_iters"])
                    perms = g[p + "perms"]
                    tr.permutation_fn = lambda ep, n, perms=perms: perms[ep].astype(np.int64)
                    info = tr.train(buf)
                    ref = dict(zip(("value_loss", "policy_loss", "dist_entropy", "actor_grad_norm", "critic_grad_norm", "ratio"),
                                   g[p + "train_info"]))
                    di = max(abs(info[k] - ref[k]) / max(1.0, abs(ref[k])) for k in ref)
                    worst = []
                    for tag, net in (("actor", pol.actor), ("critic", pol.critic)):
                        for k in net.layout:
                            key = p + tag + "." + k
                            stride = int(g[key + ":meta"][0])
                            got = net.view(k).detach().cpu().numpy().astype(np.float64).reshape(-1)[::stride]
                            refv = g[key + ":sample"].astype(np.float64)
                            d = np.abs(got - refv)
                            bad = d > 3e-6 + 2e-5 * np.abs(refv)
                            if bad.any():
                                worst.append("%s.%s %d/%d max %.2e" % (tag, k, int(bad.sum()), bad.size, d.max()))
                    print("%-28s be=%d chunk=%-3d it%d  dlogp %.1e dval %.1e dh %.1e  dinfo %.1e  bad: %s" %
                          (name, backend, chunk, it, dl, dv, dh, di, "; ".join(worst) or "-"), flush=True)
                    buf.after_update()
                pol.close()


if __name__ == "__main__":
    main()
