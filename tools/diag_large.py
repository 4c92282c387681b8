"""Diagnostic: device time of salp_step at large batch sizes, sorted by K (run on the GPU box)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from grasp_lab_salp_b200 import SalpBatch, default_params

dev = torch.device('cuda', 0)
for n in (65536, 262144, 1048576):
    b = SalpBatch(n, default_params(), seed=0)
    b.reset_device()
    g = torch.Generator(device=dev)
    g.manual_seed(1234)
    u = torch.rand((4, n, 3), generator=g, device=dev)
    u[..., 2] = u[..., 2] * 2 - 1
    for sort in (True, False):
        for i in range(3):
            b.step_device(u[i % 4], sort_by_k=sort)
        steps = 10
        st = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        en = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        sub = 0
        for i in range(steps):
            st[i].record()
            b.step_device(u[i % 4], sort_by_k=sort)
            en[i].record()
            sub += int(b.dev["substeps"].sum())
        torch.cuda.synchronize()
        ms = np.array([s.elapsed_time(e) for s, e in zip(st, en)])
        print(f"n={n:8d} sort={sort}: {ms.mean():.3f} ms/step -> {n / ms.mean() * 1e3 / 1e6:.1f} M env-steps/s, "
              f"{sub / ms.sum() * 1e3 / 1e9:.1f} G substeps/s")
    b.close()
