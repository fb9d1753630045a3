"""Launch-geometry sweep of the env step kernel (tuning aid; run on the GPU box)."""
import argparse
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dcc_b200.envs import CudaVecEnv


def time_cfg(env, acts, steps=60, warm=10):
    for t in range(warm):
        env.step(acts[t % len(acts)])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for t in range(steps):
        env.step(acts[t % len(acts)])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8)
    ap.add_argument("--m", type=int, default=64)
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--force", type=float, default=0.0)
    ap.add_argument("--generic", action="store_true")
    args = ap.parse_args()
    E, N, M = args.envs, args.n, args.m
    env = CudaVecEnv(E, N, M, comm_force_scale=args.force, reference_compat=args.force == 0.0, pos_pois="synthetic")
    acts = [torch.randn(E, N, 2, device="cuda") for _ in range(8)]
    env.reset()
    D = env.obs_dim
    balg = 4 * D + 13 + 64 + (2 * M + 5) / N
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    spec = env.use_specialized(not args.generic)
    print("specialised kernel:", spec)
    for wpc in ((4,) if spec else (1, 2, 4, 8, 16)):
        for ctas in (0, 4 * sms, 5 * sms, 8 * sms, 16 * sms, 32 * sms, (E + wpc - 1) // wpc // 4, (E + wpc - 1) // wpc // 2, (E + wpc - 1) // wpc):
            try:
                env.set_launch(wpc, ctas)
            except Exception as ex:
                print("wpc=%d ctas=%d: %s" % (wpc, ctas, ex))
                continue
            us = time_cfg(env, acts)
            print("N=%d M=%d E=%d wpc=%2d ctas=%6d : %8.1f us/step  %6.0f GB/s alg  %.3f G agent-steps/s" % (
                N, M, E, wpc, ctas, us, E * N * balg / us / 1e3, E * N / us / 1e3), flush=True)
    env.close()


if __name__ == "__main__":
    main()
