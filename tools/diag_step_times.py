"""Diagnostic: per-step device time of salp_step at small batch sizes (run on the GPU box)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from grasp_lab_salp_b200 import SalpBatch, default_params

dev = torch.device('cuda', 0)


def run(n, steps, pipeline, sort=False, A=32, seed=0):
    b = SalpBatch(n, default_params(), seed=seed)
    b.reset_device()
    g = torch.Generator(device=dev)
    g.manual_seed(1234)
    u = torch.rand((A, n, 3), generator=g, device=dev)
    u[..., 2] = u[..., 2] * 2 - 1
    for i in range(5):
        b.step_device(u[i % A], pipeline=pipeline, sort_by_k=sort)
    st = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    en = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    for i in range(steps):
        st[i].record()
        b.step_device(u[i % A], pipeline=pipeline, sort_by_k=sort)
        en[i].record()
    torch.cuda.synchronize()
    b.check()
    ms = np.array([s.elapsed_time(e) for s, e in zip(st, en)])
    print(f"n={n:6d} pipeline={pipeline} sort={sort}: mean {ms.mean():.3f} median {np.median(ms):.3f} min {ms.min():.3f} "
          f"max {ms.max():.3f} ms -> {n / ms.mean() * 1e3 / 1e6:.2f} M env-steps/s")


for n in (1024, 4096, 4736, 8192, 16384):
    run(n, 100, True)
    run(n, 100, False)
