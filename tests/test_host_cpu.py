"""CPU tests of the host-side logic around the learner kernels: config loading (the reference's YAMLs drive the build
unchanged), parameter layout / checkpoint key names, env sharding and the torch.distributed plumbing (gloo,
world_size 2) including the global-normalisation contract of the sharded PPO gradient."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_CFG = "/root/reference/uav_dcc_control/config"


def test_yaml_float_resolver_and_merge_order(tmp_path):
    from dcc_b200.utils.config import load_config
    (tmp_path / "env_config").mkdir()
    (tmp_path / "algo_config").mkdir()
    (tmp_path / "env_config" / "dcc.yaml").write_text("num_agents: 8\nppo_epoch: 15\nn_eval_rollout_threads: 16\n")
    (tmp_path / "algo_config" / "mappo.yaml").write_text("actor_lr: 5e-4\nopti_eps: 1e-5\nn_eval_rollout_threads: 1\nsave_gifs: false\n")
    (tmp_path / "expt.yaml").write_text("save_gifs: True\nload_buffer_path: None\nseed: 3\n")
    cfg = load_config(str(tmp_path))
    assert isinstance(cfg.actor_lr, float) and cfg.actor_lr == 5e-4 and cfg.opti_eps == 1e-5   # PyYAML alone gives str
    assert cfg.n_eval_rollout_threads == 1 and cfg.save_gifs is True                              # later file wins
    assert cfg.num_agents == 8 and cfg.seed == 3 and cfg.load_buffer_path is None
    assert cfg.critic_lr == 5e-4 and cfg.gamma == 0.99                                            # defaults fill the rest


@pytest.mark.skipif(not os.path.isdir(REF_CFG), reason="reference not mounted")
def test_defaults_equal_the_shipped_yaml_files():
    from dcc_b200.utils.config import DEFAULTS, load_config
    ref = vars(load_config(REF_CFG))
    new_keys = {"reference_compat", "pos_pois_path", "device", "chunk_rows", "gemm_backend"}
    for k, v in DEFAULTS.items():
        if k in new_keys:
            continue
        assert ref[k] == v, (k, ref[k], v)
    assert set(ref) - new_keys == set(DEFAULTS) - new_keys


def test_unsupported_branches_fail_loudly():
    from dcc_b200.utils.config import check_supported, load_config
    check_supported(load_config(None))
    for key, val in (("use_popart", True), ("num_mini_batch", 0), ("layer_N", 4), ("layer_N", 0)):
        cfg = load_config(None)
        setattr(cfg, key, val)
        with pytest.raises(NotImplementedError):
            check_supported(cfg)
    # recurrent policies are implemented (GRU x recurrent_N <= 4, centralised critic)
    for extra in (dict(use_recurrent_policy=True), dict(use_naive_recurrent_policy=True),
                  dict(use_recurrent_policy=True, recurrent_N=2, data_chunk_length=5)):
        cfg = load_config(None)
        for k, v in extra.items():
            setattr(cfg, k, v)
        check_supported(cfg)
    for extra in (dict(use_recurrent_policy=True, recurrent_N=5), dict(use_recurrent_policy=True, use_centralized_V=False),
                  dict(use_recurrent_policy=True, data_chunk_length=0)):
        cfg = load_config(None)
        for k, v in extra.items():
            setattr(cfg, k, v)
        with pytest.raises(NotImplementedError):
            check_supported(cfg)
    # update-path switches of mappo.yaml that ARE implemented
    for key, val in (("num_mini_batch", 2), ("use_valuenorm", False), ("use_huber_loss", False), ("use_gae", False),
                     ("use_clipped_value_loss", False), ("use_max_grad_norm", False), ("weight_decay", 1e-4),
                     ("use_proper_time_limits", True), ("use_linear_lr_decay", False), ("use_ReLU", False),
                     ("use_feature_normalization", False), ("use_centralized_V", False), ("use_orthogonal", False),
                     ("layer_N", 3), ("stacked_frames", 4), ("use_stacked_frames", True), ("use_obs_instead_of_state", True)):
        cfg = load_config(None)
        setattr(cfg, key, val)
        check_supported(cfg)


def test_net_layout_matches_reference_parameter_counts():
    from dcc_b200.algos.mappo import FC_H, TRUNK, net_layout
    lay, n = net_layout(110, 256, 2, "act.action_out.fc_mean", logstd=True)
    fc_h = 256 * 256 + 256 + 256 + 256
    assert n + fc_h == 162272                       # R_Actor at 4 UAV / 20 PoI
    layc, nc = net_layout(440, 256, 1, "v_out")
    assert nc + fc_h == 247153
    assert net_layout(338, 256, 2, "act.action_out.fc_mean", True)[1] + fc_h == 221096
    assert net_layout(2704, 256, 1, "v_out")[1] + fc_h == 831265
    assert list(lay)[:10] == list(TRUNK) and list(lay)[-1] == "act.action_out.logstd._bias" and len(FC_H) == 4
    off = 0
    for k, (o, shp) in lay.items():
        assert o == off
        off += int(np.prod(shp))
    assert off == n


def test_shard_envs_partitions_the_env_axis():
    from dcc_b200.parallel import shard_envs
    for total, world in ((524288, 8), (65536, 1), (10, 3), (7, 8)):
        spans = [shard_envs(total, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from dcc_b200.parallel import init_from_env, shard_envs
    from oracle import mappo_oracle as mo
    comm = init_from_env(backend="gloo")
    assert (comm.world, comm.rank) == (world, rank)
    # the sharded-gradient contract of dcc_mappo_epoch_grads: each rank sums per-row gradients of ITS env shard
    # already divided by the GLOBAL row count; a SUM all-reduce then equals the single-process big-batch gradient.
    rng = np.random.default_rng(0)
    E, T, D, H = 6, 5, 7, 8
    x = rng.normal(size=(T, E, D)); dv = rng.normal(size=(T, E, 1))
    p = {"base.feature_norm.weight": np.ones(D) + 0.1 * rng.normal(size=D), "base.feature_norm.bias": 0.1 * rng.normal(size=D),
         "base.mlp.fc1.0.weight": rng.normal(size=(H, D)), "base.mlp.fc1.0.bias": rng.normal(size=H),
         "base.mlp.fc1.2.weight": np.ones(H), "base.mlp.fc1.2.bias": np.zeros(H),
         "base.mlp.fc2.0.0.weight": rng.normal(size=(H, H)), "base.mlp.fc2.0.0.bias": rng.normal(size=H),
         "base.mlp.fc2.0.2.weight": np.ones(H), "base.mlp.fc2.0.2.bias": np.zeros(H),
         "v_out.weight": rng.normal(size=(1, H)), "v_out.bias": np.zeros(1)}
    net = mo.make_critic(p)
    net.forward(x.reshape(T * E, D))
    full = net.backward(dv.reshape(T * E, 1) / (T * E))
    lo, hi = shard_envs(E, world, rank)
    net.forward(x[:, lo:hi].reshape(-1, D))
    part = net.backward(dv[:, lo:hi].reshape(-1, 1) / (T * E))          # divided by the GLOBAL count
    flat = torch.from_numpy(np.concatenate([part[k].reshape(-1) for k in p]))
    comm.all_reduce_sum_(flat)
    ref = np.concatenate([full[k].reshape(-1) for k in p])
    ok = bool(np.allclose(flat.numpy(), ref, rtol=1e-10, atol=1e-12)) and comm.calls == 1
    t = torch.tensor([float(rank + 1)])
    comm.broadcast_(t, src=0)
    ok = ok and float(t) == 1.0
    comm.barrier()
    q.put((rank, ok))
    dist.destroy_process_group()


def test_gloo_world2_sharded_gradient_allreduce():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]


def test_comm_is_a_noop_without_a_process_group():
    import torch
    from dcc_b200.parallel import Comm
    c = Comm()
    t = torch.ones(3)
    assert (c.world, c.rank) == (1, 0) and c.all_reduce_sum_(t) is t and c.calls == 0


def test_headless_render_and_trajectory_recorder(tmp_path):
    """f-3: the pyglet viewer's replacement rasterises from the compact state; the recorder round-trips."""
    from dcc_b200.envs.headless_render import TrajectoryRecorder, rasterize
    rng = np.random.default_rng(0)
    N, M = 4, 20
    poi = rng.uniform(-1, 1, (M, 2))
    pos = np.array([[0.0, 0.0], [0.3, 0.0], [1.5, 1.5], [-1.0, 0.5]])
    energy = np.zeros(M, np.uint8); energy[:5] = 5; energy[5:8] = 2
    adj = np.array([0b0010, 0b0001, 0, 0], np.uint32)
    f = rasterize(pos, poi, energy, adj, size=200)
    assert f.shape == (200, 200, 3) and f.dtype == np.uint8
    assert (f != 255).any() and (f == 255).all(-1).mean() > 0.5          # something drawn on a white canvas
    # the comm line between UAV 0 and 1 (blue) is present; without adjacency it is not
    blue = lambda im: int(((im[..., 2] > 180) & (im[..., 0] < 80)).sum())
    assert blue(f) > 0 and blue(rasterize(pos, poi, energy, np.zeros(4, np.uint32), size=200)) == 0
    rec = TrajectoryRecorder(poi, 0.2, 0.4)
    for t in range(3):
        pv = np.zeros((1, N, 4)); pv[0, :, :2] = pos + 0.01 * t
        rec.add(pv, energy[None], np.array([t & 1], np.uint8), adj[None], np.array([0.25]), np.array([-1.0]))
    assert abs(rec.connectivity_rate() - 1.0 / 3.0) < 1e-12
    path = str(tmp_path / "traj.npz")
    rec.save(path)
    z = np.load(path)
    assert z["pos_vel"].shape == (3, 1, N, 4) and z["adj"].shape == (3, 1, N) and z["poi_xy"].shape == (M, 2)
    rec.save_gif(str(tmp_path / "t.gif"), size=64)
    assert os.path.getsize(str(tmp_path / "t.gif")) > 100


@pytest.mark.parametrize("name", ["ship_4x20", "xavier_tanh_nofn_decv", "rnn2_3x20", "rnn1_xavier_4x20"])
def test_reference_order_initialisation(name):
    """MAPPOPolicy draws its initial weights with the reference's torch RNG consumption order (actor trunk, actor head,
    critic trunk, critic head; orthogonal / xavier_uniform; ReLU / tanh gain): same seed => the reference's weights.
    Goldens: tests/golden/make_golden_mappo.py::run_init_case (unmodified reference Learner, seeded by its own
    seed_everything)."""
    import json
    import torch
    from dcc_b200.algos.mappo import _reference_init
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "init_%s.npz" % name))
    c = json.loads(str(z["cfg"]))
    D, Hd, N = c["obs_dim"], c["hidden"], c["n_agents"]
    S = N * D if c["use_centralized_V"] else D
    kw = dict(use_orthogonal=c["use_orthogonal"], use_relu=c["use_ReLU"], feature_norm=c["use_feature_normalization"],
              layer_N=c.get("layer_N", 1), recurrent_N=c.get("recurrent_N", 0))
    torch.manual_seed(c["seed"])
    sd_a, head_a = _reference_init(D, Hd, 2, c["gain"], **kw)
    sd_a["act.action_out.fc_mean.weight"], sd_a["act.action_out.fc_mean.bias"] = head_a.weight.data, head_a.bias.data
    sd_a["act.action_out.logstd._bias"] = torch.zeros(2, 1)
    sd_c, head_c = _reference_init(S, Hd, 1, 1.0, **kw)
    sd_c["v_out.weight"], sd_c["v_out.bias"] = head_c.weight.data, head_c.bias.data
    n_checked = 0
    for tag, sd in (("actor", sd_a), ("critic", sd_c)):
        ref_keys = {k[len(tag) + 1:-7] for k in z.files if k.startswith(tag + ".") and k.endswith(":sample")}
        assert ref_keys == set(sd.keys()), (tag, ref_keys ^ set(sd.keys()))
        for k, v in sd.items():
            flat = v.numpy().astype(np.float64).reshape(-1)
            stride, s, ss = z["%s.%s:meta" % (tag, k)]
            # same RNG stream; the QR inside orthogonal_ rounds differently with a different BLAS thread count (~1e-7)
            assert np.allclose(flat[::int(stride)], z["%s.%s:sample" % (tag, k)], rtol=0, atol=2e-6), (tag, k)
            assert abs(flat.sum() - s) <= 1e-5 * max(1.0, np.abs(flat).sum()) and abs((flat ** 2).sum() - ss) <= 1e-5 * max(1.0, ss)
            n_checked += 1
    assert n_checked >= 26


def test_fp16_split_emulation():
    """Numerics behind the fp16 hi/lo split GEMMs (csrc/dcc_tc.cuh, tc_gemm_fwd_kernel<true>), emulated in NumPy with
    float64 products of the rounded halves: x = hi + lo with hi = fp16(x), lo = fp16(x - hi), product = hi*hi + hi*lo +
    lo*hi.  With the weight image pre-scaled by 2^8 (TC_F16_WSCALE) the split is as accurate as the 3xTF32 split; an
    unscaled image is not (the lo halves of |w| ~ 0.05 are fp16-subnormal).  Second half: the next-round plan for the
    weight gradients — ONE power-of-two scale per dZ tensor keeps 3xTF32 accuracy over 3.5 decades of row scales."""
    rng = np.random.default_rng(0)
    f = lambda a: a.astype(np.float64)   # noqa: E731

    def split16(a):
        h = a.astype(np.float16)
        return h, (a - h.astype(np.float32)).astype(np.float16)

    def split_tf32(a):
        rna = lambda v: ((v.view(np.uint32) + 0x1000) & 0xFFFFE000).view(np.float32)   # noqa: E731
        h = rna(a.copy())
        return h, rna((a - h).astype(np.float32))

    def prod(ah, al, bh, bl):
        return f(ah) @ f(bh).T + f(ah) @ f(bl).T + f(al) @ f(bh).T

    M, K, N = 256, 352, 256
    x = rng.normal(0, 1, (M, K)).astype(np.float32)            # a LayerNorm output
    x[:, :40] *= 1e-3
    W = (rng.normal(0, 1, (N, K)) * np.sqrt(2.0 / K)).astype(np.float32)
    ref = f(x) @ f(W).T
    rms = lambda y, r: np.sqrt(np.mean((y - r) ** 2)) / np.sqrt(np.mean(r ** 2))   # noqa: E731
    e_tf32 = rms(prod(*split_tf32(x), *split_tf32(W)), ref)
    e_f16_scaled = rms(prod(*split16(x), *split16(W * np.float32(256.0))) / 256.0, ref)
    e_f16_plain = rms(prod(*split16(x), *split16(W)), ref)
    assert e_tf32 < 1e-7 and e_f16_scaled < 1e-7 and e_f16_scaled < 1.1 * e_tf32
    assert e_f16_plain > 2 * e_f16_scaled
    # weight-gradient shape: G = dZ^T X, rows of dZ spread over 3.5 decades, half of the entries masked by ReLU
    R = 2048
    X = rng.normal(0, 1, (R, K)).astype(np.float32)
    dz = (rng.normal(0, 1, (R, N)) * 10 ** rng.uniform(-10, -6.5, (R, 1))).astype(np.float32)
    dz[rng.random((R, N)) < 0.5] = 0
    refg = f(dz).T @ f(X)
    dzh, dzl = split_tf32(dz)
    xh, xl = split_tf32(X)
    e_tf32 = rms(f(dzh).T @ f(xh) + f(dzh).T @ f(xl) + f(dzl).T @ f(xh), refg)
    S = np.float32(2.0 ** (14 - np.ceil(np.log2(np.abs(dz).max()))))
    dh, dl = split16(dz * S)
    xh, xl = split16(X)
    assert np.isfinite(dh.astype(np.float32)).all()
    e_f16 = rms((f(dh).T @ f(xh) + f(dh).T @ f(xl) + f(dl).T @ f(xh)) / float(S), refg)
    dh0, dl0 = split16(dz)                                   # unscaled: the gradients vanish into fp16 subnormals
    e_f16_plain = rms(f(dh0).T @ f(xh) + f(dh0).T @ f(xl) + f(dl0).T @ f(xh), refg)
    assert e_tf32 < 1e-7 and e_f16 < 1.1 * e_tf32 and e_f16_plain > 1e-3


def test_load_checkpoint_reads_a_reference_written_agent_pkl(tmp_path):
    """MAPPOTrainer.load_model must read the reference's own checkpoint: `pickle.dump(self.policy, f)`
    (algos/mappo.py:237-240).  The fixture was written by the UNMODIFIED reference (tests/golden/make_golden_mappo.py
    run_pickle_case); it is read here WITHOUT the reference on sys.path, through the restricted unpickler."""
    import pickle
    import torch
    from dcc_b200.algos.mappo import CHECKPOINT_FORMAT, _CheckpointUnpickler, load_checkpoint
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_agent_3x20_h32.npz"))
    sd = load_checkpoint(os.path.join(ROOT, "tests", "golden", "ref_agent_3x20_h32.pkl"))
    assert sd["format"] == "reference.policy.pickle"
    n = 0
    for tag in ("actor", "critic"):
        keys = [k[len(tag) + 1:] for k in g.files if k.startswith(tag + ".")]
        assert list(sd[tag].keys()) == keys            # torch's state_dict order, fc_h block included
        for k in keys:
            assert np.array_equal(sd[tag][k].numpy(), g[tag + "." + k]), (tag, k)
            n += 1
    assert n == 33    # actor 17 (incl. fc_h and logstd), critic 16
    # a recurrent policy written by the reference (R_Actor / R_Critic with an RNNLayer: nn.GRU x 2 + LayerNorm, rnn.py:8-22)
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_agent_rnn2_3x20_h32.npz"))
    sd = load_checkpoint(os.path.join(ROOT, "tests", "golden", "ref_agent_rnn2_3x20_h32.pkl"))
    for tag in ("actor", "critic"):
        keys = [k[len(tag) + 1:] for k in g.files if k.startswith(tag + ".")]
        assert list(sd[tag].keys()) == keys and "rnn.rnn.weight_hh_l1" in keys and "rnn.norm.bias" in keys
        for k in keys:
            assert np.array_equal(sd[tag][k].numpy(), g[tag + "." + k]), (tag, k)
    # this build's own format round-trips through the same reader; anything else is refused with a clear error
    own = {"actor": {"w": torch.ones(2)}, "critic": {"w": torch.zeros(2)}, "format": CHECKPOINT_FORMAT}
    with open(tmp_path / "agent.pkl", "wb") as f:
        pickle.dump(own, f)
    back = load_checkpoint(str(tmp_path / "agent.pkl"))
    assert back["format"] == CHECKPOINT_FORMAT and torch.equal(back["actor"]["w"], torch.ones(2))
    with open(tmp_path / "bad.pkl", "wb") as f:
        pickle.dump([1, 2, 3], f)
    with pytest.raises(ValueError):
        load_checkpoint(str(tmp_path / "bad.pkl"))
    import io

    class Evil:
        def __reduce__(self):
            return (eval, ("1+1",))
    with pytest.raises(pickle.UnpicklingError):
        _CheckpointUnpickler(io.BytesIO(pickle.dumps(Evil()))).load()
