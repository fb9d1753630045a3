"""GPU parity tests of the recurrent policies (use_recurrent_policy / use_naive_recurrent_policy — a GRU with
recurrent_N layers and a LayerNorm between the trunk and the head of both nets, algos/algo_utils/rnn.py), through the C ABI
(dcc_mappo_act_rnn, dcc_mappo_seq_grads).

Checked against golden vectors recorded from the UNMODIFIED reference learner with the switch flipped
(tests/golden/mappo_rnn_*.npz: chunked BPTT, whole-episode BPTT, two GRU layers x two minibatches with chunks that
straddle sequences, hidden 256 in front of the tcgen05 trunk).  Same tolerances as tests/test_mappo_cuda.py.
"""
import numpy as np
import pytest

from test_mappo_cuda import ALL_CASES, BACKENDS, build, check_params, fill_buffer, load

pytestmark = pytest.mark.gpu

RNN_CASES = [n for n in ALL_CASES if n.startswith("rnn_")]


def fill_rnn(buf, g, p):
    import torch
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(buf.device)   # noqa: E731
    hs_c = g[p + "rnn_states_critic"]
    assert np.array_equal(hs_c, np.broadcast_to(hs_c[:, :, :1], hs_c.shape))     # identical for the agents of an env
    buf.rnn_a.copy_(t(g[p + "rnn_states"]))
    buf.rnn_c.copy_(t(hs_c[:, :, 0]))


@pytest.mark.parametrize("chunk_rows", [0, 40])
@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("name", RNN_CASES)
def test_recurrent_learner_vs_reference_golden(name, backend, chunk_rows):
    """Teacher-forced replay of two reference iterations with a recurrent policy: per rollout step the recorded hidden
    states go in and log-probs / values / the next hidden states must come out; then GAE and the whole update (the
    sequence chunks are replayed in the order the reference's generator drew them).  chunk_rows = 40 forces several
    passes of sequences per minibatch (and chunk < T: skipped for whole-episode BPTT, which needs a pass >= T rows)."""
    import torch
    g = load(name)
    c = g["cfg"]
    N, D, R, Hd = c["n_agents"], c["obs_dim"], c["recurrent_N"], c["hidden"]
    T, E = g["it1_actions"].shape[:2]
    L = c["data_chunk_length"] if c["use_recurrent_policy"] else T
    if chunk_rows and chunk_rows < L:
        pytest.skip("a pass must hold one whole sequence")
    cfg, pol, tr, buf = build(c, E, T, gemm_backend=backend, chunk_rows=chunk_rows)
    assert pol.recurrent_N == R and buf.recurrent
    dev = buf.device
    for it in range(1, c["iters"] + 1):
        p = "it%d_" % it
        fill_buffer(buf, g, p)
        fill_rnn(buf, g, p)
        vn = tr.value_normalizer.state.cpu().numpy()[:3]
        assert np.allclose(vn, g[p + "vn_before"], rtol=1e-5, atol=1e-12)
        ftol = 1e-5 if it == 1 else 1e-4
        logp = torch.empty(E * N, device=dev)
        vals = torch.empty(E, device=dev)
        for t in range(T + 1):
            if t < T:
                ha, hc = pol._act_rnn(buf.obs[t].contiguous(), E, buf.rnn_a[t], buf.rnn_c[t], buf.masks_te[t], 1, False,
                                      buf.actions[t].contiguous(), logp, vals)
                assert np.allclose(logp.cpu().numpy().reshape(E, N, 1), g[p + "logp"][t], rtol=ftol, atol=ftol), t
                keep = g[p + "masks"][t + 1].reshape(E * N, 1, 1)
                assert np.allclose(ha.cpu().numpy() * keep, g[p + "rnn_states"][t + 1].reshape(E * N, R, Hd), rtol=ftol, atol=ftol), t
                assert np.allclose(hc.cpu().numpy() * keep, g[p + "rnn_states_critic"][t + 1].reshape(E * N, R, Hd), rtol=ftol, atol=ftol), t
            else:       # the bootstrap value (learner.py:278-287): get_values with the reference's per-agent layouts
                v = pol.get_values(buf.share_obs[t].reshape(E * N, N * D), buf.rnn_states_critic[t].reshape(E * N, R, Hd),
                                   buf.masks[t].reshape(E * N, 1), rows_repeated=True)
                vals = v.reshape(E, N)[:, 0]
            assert np.allclose(vals.cpu().numpy().reshape(E, 1, 1), g[p + "value_preds"][t][:, :1], rtol=ftol, atol=ftol), t
        buf.compute_returns(None, tr.value_normalizer, policy=pol)
        ret = buf.returns_te.cpu().numpy()[:-1]
        ref_all = g[p + "returns"][:, :, 0, 0]
        assert np.allclose(ret, ref_all[:-1], rtol=1e-5, atol=1e-4)
        buf.returns_te.copy_(torch.from_numpy(np.ascontiguousarray(ref_all)).to(dev))
        pol.lr_decay(it, c["n_iters"])
        assert abs(pol.lr_actor_now - float(g[p + "lr"])) < 1e-12
        perms = g[p + "perms"]
        tr.permutation_fn = lambda ep, n, perms=perms: perms[ep].astype(np.int64)
        info = tr.train(buf)
        ref = dict(zip(("value_loss", "policy_loss", "dist_entropy", "actor_grad_norm", "critic_grad_norm", "ratio"),
                       g[p + "train_info"]))
        for k in ref:
            assert abs(info[k] - ref[k]) <= 5e-5 * max(1.0, abs(ref[k])), (it, k, info[k], ref[k])
        vn = tr.value_normalizer.state.cpu().numpy()[:3]
        assert np.allclose(vn, g[p + "vn_after"], rtol=1e-5, atol=1e-12)
        frac = 0.02 if c.get("num_mini_batch", 1) > 1 else 0.0
        check_params("actor it%d" % it, pol.actor, g, p + "actor.", max_bad_frac=frac)
        check_params("critic it%d" % it, pol.critic, g, p + "critic.", max_bad_frac=frac)
        buf.after_update()


@pytest.mark.parametrize("over", [dict(use_recurrent_policy=True), dict(use_naive_recurrent_policy=True),
                                  dict(use_recurrent_policy=True, recurrent_N=2, data_chunk_length=7, num_mini_batch=2)])
def test_recurrent_learner_end_to_end(over):
    """The re-hosted Learner with the recurrent switches of mappo.yaml flipped: rollout (hidden states written to the
    buffer by the policy kernels, zeroed where episodes end, carried into the next rollout by after_update) and update
    run for two iterations; MLP entry points refuse a recurrent handle."""
    import torch
    from dcc_b200 import _lib
    from dcc_b200.learner import Learner
    from dcc_b200.utils.config import load_config
    cfg = load_config(None, num_agents=3, num_pois=10, n_rollout_threads=32, max_ep_len=21, ppo_epoch=2, n_iters=3,
                      algo_hidden_size=64, n_eval_rollout_threads=0, n_render_rollout_threads=0, save_model=False, **over)
    lr = Learner(cfg)
    assert not lr.compact and lr.rl_buffer.recurrent
    before = lr.policy.actor.view("rnn.rnn.weight_hh_l0").clone()
    last = None
    for it in range(1, 3):
        lr.policy.lr_decay(it, cfg.n_iters)
        lr.rollout(lr.rl_buffer, lr.train_envs)
        buf = lr.rl_buffer
        assert torch.isfinite(buf.rnn_a).all() and torch.isfinite(buf.rnn_c).all()
        assert float(buf.rnn_a[1:].abs().max()) > 0 and float(buf.rnn_c[1:].abs().max()) > 0
        if last is not None:       # slot 0 = the previous rollout's final states (shared_buffer.py:146-147)
            assert torch.equal(buf.rnn_a[0], last[0]) and torch.equal(buf.rnn_c[0], last[1])
        done = buf.masks_te[1:] == 0
        if bool(done.any()):
            assert float(buf.rnn_c[1:][done].abs().max()) == 0.0
        last = (buf.rnn_a[-1].clone(), buf.rnn_c[-1].clone())
        info = lr.rl_update()
        assert all(np.isfinite(v) for v in info.values()), info
    assert not torch.equal(before, lr.policy.actor.view("rnn.rnn.weight_hh_l0"))
    with pytest.raises(_lib.DccError):
        lr.policy.lib and _lib.check(lr.policy.lib.dcc_mappo_act(lr.policy._h, lr.policy._ptr(lr.policy.actor.params), None,
                                                                 lr.policy._ptr(lr.rl_buffer.obs[0]), 32, 0, 0, 0,
                                                                 lr.policy._ptr(lr.rl_buffer.actions[0]), None, None,
                                                                 lr.policy._stream()), "dcc_mappo_act")
    sd = lr.policy.state_dict()
    assert "rnn.rnn.weight_ih_l0" in sd["actor"] and "rnn.norm.bias" in sd["critic"]


def test_reference_recurrent_checkpoint_loads_and_acts(tmp_path):
    """An agent.pkl written by the UNMODIFIED reference with use_recurrent_policy / recurrent_N = 2 (pickled MAPPOPolicy object,
    mappo.py:237-240) loads into the recurrent MAPPOPolicy; one step of get_actions / get_values / evaluate_actions on it equals
    the float64 oracle run on the same parameters; save_model -> load_model round-trips the GRU tensors."""
    import os
    import torch
    from oracle import mappo_oracle as mo
    from dcc_b200.algos import MAPPOPolicy, MAPPOTrainer
    from dcc_b200.envs.spaces import Box
    from dcc_b200.utils.config import load_config
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    g = np.load(os.path.join(here, "ref_agent_rnn2_3x20_h32.npz"))
    N, M, Hd, R = 3, 20, 32, 2
    D = 4 + 2 * (N - 1) + 5 * M
    cfg = load_config(None, num_agents=N, num_pois=M, algo_hidden_size=Hd, use_recurrent_policy=True, recurrent_N=R,
                      n_rollout_threads=8, n_eval_rollout_threads=0)
    pol = MAPPOPolicy(cfg, Box(-np.inf, np.inf, (D,)), Box(-np.inf, np.inf, (N * D,)), Box(-1, 1, (2,)))
    tr = MAPPOTrainer(cfg, pol)
    tr.load_model(os.path.join(here, "ref_agent_rnn2_3x20_h32.pkl"))
    ap = {k[6:]: g[k] for k in g.files if k.startswith("actor.") and "fc_h" not in k}
    cp = {k[7:]: g[k] for k in g.files if k.startswith("critic.") and "fc_h" not in k}
    for k in pol.actor.layout:
        assert np.array_equal(pol.actor.view(k).cpu().numpy().reshape(-1), ap[k].reshape(-1)), k
    rng = np.random.default_rng(3)
    E = 8
    obs = rng.normal(0, 1, (E, N, D)).astype(np.float32)
    ha = rng.normal(0, 0.5, (E * N, R, Hd)).astype(np.float32)
    hc = np.repeat(rng.normal(0, 0.5, (E, 1, R, Hd)).astype(np.float32), N, axis=1).reshape(E * N, R, Hd)
    masks = (rng.random((E, 1)) > 0.3).astype(np.float32).repeat(N, 1).reshape(E * N, 1)
    act = rng.normal(0, 1, (E * N, 2)).astype(np.float32)
    t = lambda a: torch.from_numpy(a).cuda()   # noqa: E731
    v, lp, ent = pol.evaluate_actions(None, t(obs), t(ha), t(hc), t(act), t(masks))
    oa, oc = mo.RecurrentNet(mo.make_actor(ap)), mo.RecurrentNet(mo.make_critic(cp))
    mean = oa.forward(obs.reshape(E * N, D), ha, masks, 1)
    lp_ref, ent_ref = mo.gaussian_logp_entropy(mean, ap["act.action_out.logstd._bias"].reshape(1, -1).astype(np.float64), act)
    sx = np.repeat(obs.reshape(E, 1, N * D), N, axis=1).reshape(E * N, N * D)
    v_ref = oc.forward(sx, hc, masks, 1)
    assert np.allclose(lp.cpu().numpy(), lp_ref, rtol=1e-5, atol=1e-5)
    assert np.allclose(v.cpu().numpy(), v_ref, rtol=1e-5, atol=1e-5)
    assert abs(float(ent) - float(ent_ref)) < 1e-6
    # deterministic get_actions returns the mean and the new hidden states of both nets
    vals, a_det, _, ha2, hc2 = pol.get_actions(None, t(obs), t(ha), t(hc), t(masks), deterministic=True)
    assert np.allclose(a_det.cpu().numpy(), mean, rtol=1e-5, atol=1e-5)
    assert np.allclose(ha2.cpu().numpy(), oa.h_final, rtol=1e-5, atol=1e-5) and np.allclose(hc2.cpu().numpy(), oc.h_final, rtol=1e-5, atol=1e-5)
    a_only, ha3 = pol.act(t(obs), t(ha), t(masks), deterministic=True)
    assert torch.allclose(a_only.view(-1, 2), a_det) and torch.allclose(ha3, ha2)
    # this build's checkpoint format carries the GRU tensors too
    tr.save_model(str(tmp_path))
    pol2 = MAPPOPolicy(cfg, Box(-np.inf, np.inf, (D,)), Box(-np.inf, np.inf, (N * D,)), Box(-1, 1, (2,)))
    MAPPOTrainer(cfg, pol2).load_model(str(tmp_path))
    assert torch.equal(pol2.actor.params, pol.actor.params) and torch.equal(pol2.critic.params, pol.critic.params)
