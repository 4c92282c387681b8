"""GPU diagnostic: per-step device time of salp_step below the K-sort crossover -- uniform-K batches
(fixed cost, shape-moving substeps, coast substeps) and uniform-random actions, for the kernel the
launcher picks by default, the fused kernel and the K-sorted variants.
    SALP_PIPE_VARIANT=3 python tools/diag_small_batch.py     # round-1 three-warp pipeline kernel
    python tools/diag_small_batch.py                         # four-warp pipeline kernel
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from grasp_lab_salp_b200 import SalpBatch, default_params

dev = torch.device("cuda", 0)
NS = [int(x) for x in os.environ.get("DIAG_NS", "1024,4096,4736,8192,9472,16384").split(",")]
STEPS = int(os.environ.get("DIAG_STEPS", "60"))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > L2 (126 MB)


def timed(b, actions, **kw):
    """actions: [A, n, 3]; median / mean ms per step, L2 flushed between steps."""
    A = actions.shape[0]
    for i in range(6):
        b.step_device(actions[i % A], **kw)
    st = [torch.cuda.Event(enable_timing=True) for _ in range(STEPS)]
    en = [torch.cuda.Event(enable_timing=True) for _ in range(STEPS)]
    for i in range(STEPS):
        flush.zero_()
        st[i].record()
        b.step_device(actions[i % A], **kw)
        en[i].record()
    torch.cuda.synchronize()
    b.check()
    ms = np.array([s.elapsed_time(e) for s, e in zip(st, en)])
    return float(np.median(ms)), float(ms.mean()), b.last_step_kernel


print("variant", os.environ.get("SALP_PIPE_VARIANT", "4 (default)"))
for n in NS:
    b = SalpBatch(n, default_params(), seed=0)
    b.reset_device()
    g = torch.Generator(device=dev)
    g.manual_seed(1234)
    u = torch.rand((16, n, 3), generator=g, device=dev)
    u[..., 2] = u[..., 2] * 2 - 1
    z = torch.zeros((1, n, 3), device=dev)
    cases = [("K=0", z.clone())]
    a = z.clone(); a[..., 0] = 0.5
    cases.append(("a0=.5 no coast (K=220, shape moves)", a))
    a = z.clone(); a[..., 0] = 1.0; a[..., 1] = 1.0
    cases.append(("a0=1 coast 10 s (K=1348)", a))
    cases.append(("uniform random", u))
    for label, acts in cases:
        row = []
        for kw_label, kw in (("default", {}), ("fused", dict(pipeline=False)), ("default+sort", dict(sort_by_k=True))):
            med, mean, kern = timed(b, acts, **kw)
            row.append(f"{kw_label}: {med * 1e3:7.1f} us [{kern}]")
        print(f"n={n:6d} {label:38s} " + "  ".join(row), flush=True)
    b.close()
