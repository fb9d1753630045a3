"""Runtime-shape (generic) env kernel: us/step for a few shapes and launch geometries (tuning aid)."""
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from dcc_b200.envs import CudaVecEnv
for (N, M, E) in ((8, 64, 65536), (16, 256, 32768), (6, 41, 65536)):
    for wpc in (2, 4, 8):
        env = CudaVecEnv(E, N, M, comm_force_scale=1.0, reference_compat=False, pos_pois="synthetic")
        env.use_specialized(False)
        try:
            env.set_launch(wpc, 0)
        except Exception as ex:
            print(N, M, "wpc", wpc, "refused"); env.close(); continue
        acts = [torch.randn(E, N, 2, device="cuda") for _ in range(4)]
        env.reset()
        for t in range(20):
            env.step(acts[t % 4])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(60):
            env.step(acts[t % 4])
        e1.record(); torch.cuda.synchronize()
        print("N=%d M=%d E=%d wpc=%d: %.1f us/step" % (N, M, E, wpc, e0.elapsed_time(e1) * 1e3 / 60), flush=True)
        env.close()
