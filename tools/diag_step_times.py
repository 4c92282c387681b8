"""Diagnostic: per-step device time of salp_step at small batch sizes (run on the GPU box)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from grasp_lab_salp_b200 import SalpBatch, default_params

dev = torch.device('cuda', 0)


def run(n, steps, A, seed=0, every=20):
    b = SalpBatch(n, default_params(), seed=seed)
    b.reset_device()
    g = torch.Generator(device=dev)
    g.manual_seed(1234)
    u = torch.rand((A, n, 3), generator=g, device=dev)
    u[..., 2] = u[..., 2] * 2 - 1
    st = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    en = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    stats = []
    for i in range(steps):
        st[i].record()
        b.step_device(u[i % A])
        en[i].record()
        if i % every == 0:
            ex, ey = b.state_tensor("euler_x"), b.state_tensor("euler_y")
            stats.append((i, float(ex.abs().max()), float(ey.abs().max()), float(b.state_tensor("euler_z").abs().max()),
                          int(b.state_tensor("cycle").max()), float(b.state_tensor("vel_x").abs().max()),
                          int(b.dev["substeps"].max()), float(b.dev["substeps"].float().mean())))
    torch.cuda.synchronize()
    ms = np.array([s.elapsed_time(e) for s, e in zip(st, en)])
    for (i, ex, ey, ez, cyc, vx, km, kmean) in stats:
        print(f"  step {i:4d}: {ms[i]:.3f} ms  max|roll| {ex:.3g} max|pitch| {ey:.3g} max|yaw| {ez:.3g} max cycle {cyc} max|vx| {vx:.3g} Kmax {km} Kmean {kmean:.0f}")
    print(f"n={n}: mean {ms.mean():.3f} median {np.median(ms):.3f}")


run(4096, 400, 32)
run(4096, 60, 32, seed=5)
