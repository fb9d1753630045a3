import sys, os
sys.path.insert(0, os.getcwd())
import torch
from dcc_b200.envs import CudaVecEnv
E, N, M = 32768, 16, 256
env = CudaVecEnv(E, N, M, comm_r_scale=0.95, comm_force_scale=1.0, reference_compat=False, pos_pois="synthetic")
acts = [torch.randn(E, N, 2, device="cuda") for _ in range(4)]
env.reset()
for t in range(30):
    env.step(acts[t % 4])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for t in range(40):
    env.step(acts[t % 4])
e1.record(); torch.cuda.synchronize()
print("us/step", e0.elapsed_time(e1) * 1e3 / 40)
