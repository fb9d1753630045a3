import sys, os
sys.path.insert(0, os.getcwd())
import torch
from dcc_b200.envs import CudaVecEnv
E, N, M = 65536, 8, 64
env = CudaVecEnv(E, N, M, reference_compat=True, pos_pois="synthetic")
acts = [torch.randn(E, N, 2, device="cuda") for _ in range(8)]
env.reset()
for t in range(30):
    env.step(acts[t % 8])
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(150):
        env.step(acts[t % 8])
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) * 1e3 / 150)
print("us/step %.2f" % best)
