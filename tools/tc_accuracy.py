"""Accuracy probe of the tcgen05 3xTF32 GEMM vs float64 (and vs the SIMT fp32 kernel): error growth with K and its sign
(round-toward-zero accumulation shows up as a bias that shrinks |C|)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from dcc_b200 import _lib
from test_mappo_cuda import build

c = dict(n_agents=8, n_pois=64, hidden=256, obs_dim=338, ppo_epoch=1, seed=0, n_iters=1, actor_seed=1, critic_seed=2)
cfg, pol, tr, buf = build(c, 2, 2)
lib = pol.lib
rng = np.random.default_rng(0)
for K in (32, 64, 128, 256, 352, 1024, 2720):
    M, N = 2048, 256
    A = rng.normal(0, 1, (M, K)).astype(np.float32)
    B = (rng.normal(0, 1, (N, K)) / np.sqrt(K)).astype(np.float32)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    out = {}
    for be in (1, 2, 3):
        dC = torch.zeros((M, N), dtype=torch.float32, device="cuda")
        _lib.check(lib.dcc_op_gemm(pol._h, be, 0, 1, M, N, K, dA.data_ptr(), K, dB.data_ptr(), K, dC.data_ptr(), N, 0, None), "gemm")
        torch.cuda.synchronize()
        e = dC.cpu().numpy().astype(np.float64) - ref
        out[be] = (np.abs(e).max(), np.sqrt((e ** 2).mean()), (e * np.sign(ref)).mean())
    print("K=%5d  simt: max %.2e rms %.2e bias %+.2e | 3xtf32: max %.2e rms %.2e bias %+.2e | fp16 split: max %.2e rms %.2e bias %+.2e   (|C| rms %.2f)" %
          ((K,) + out[1] + out[2] + out[3] + (np.sqrt((ref ** 2).mean()),)))
