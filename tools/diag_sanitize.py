"""GPU diagnostic for compute-sanitizer (memcheck / racecheck): a few steps through every kernel path
at small sizes -- pipeline kernel (ragged block, a K = 0 warp, natural and K-sorted order), fused
kernel natural and K-sorted, F64, randomised, both host transports (page-locked zero-copy and
pageable staged), device face, reset, trace.
    compute-sanitizer --tool racecheck python tools/diag_sanitize.py
    compute-sanitizer --tool memcheck  python tools/diag_sanitize.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grasp_lab_salp_b200 import PRECISION_F64, SalpBatch, default_params

rng = np.random.default_rng(0)
steps = int(os.environ.get("DIAG_STEPS", "2"))


def actions(n):
    a = rng.uniform([0, 0, -1], [1, 1, 1], size=(n, 3)).astype(np.float32)
    a[:8] = [[0, 0, 0], [1, 1, 1], [1, 0, -1], [0.088, 0, 0.5], [0.5, 0.5, 1e-4], [0.09, 0, 0], [0, 1, 1], [1, 1, 0]]
    if n >= 96:
        a[64:96] = 0.0          # a whole warp with K = 0
    return a


for label, n, kw, params, pinned in [
    ("pipeline, ragged block + K=0 warp", 200, dict(pipeline=True), default_params(), True),
    ("pipeline, K-sorted", 200, dict(pipeline=True, sort_by_k=True), default_params(), True),
    ("pipeline, pageable host buffers (staged)", 100, dict(pipeline=True), default_params(), False),
    ("fused", 200, dict(pipeline=False), default_params(), True),
    ("fused K-sorted", 200, dict(pipeline=False, sort_by_k=True), default_params(), True),
    ("f64", 96, dict(), default_params(precision=PRECISION_F64), True),
    ("randomised", 96, dict(), default_params(randomization=31), True),
    ("5 obstacles", 70, dict(), default_params(num_obstacles=5), True),
]:
    b = SalpBatch(n, params, seed=1)
    if not pinned:
        for name in ("obs", "terminal_obs", "reward", "terminated", "truncated"):
            setattr(b, name, np.zeros_like(getattr(b, name)))
    b.reset()
    for t in range(steps):
        b.step(actions(n), auto_reset=True, **kw)
    b.check()
    kern = b.last_step_kernel
    mask = np.zeros(n, np.uint8)
    mask[::3] = 1
    b.reset(mask)
    b.trace_cycle(1, [0.5, 0.1, 0.3])
    print(f"{label:45s} ok  [{kern}]  mean reward {float(b.reward.mean()):.3f}", flush=True)
    b.close()

import torch  # noqa: E402
b = SalpBatch(300, default_params(), seed=2)
b.reset_device()
for t in range(steps):
    b.step_device(torch.from_numpy(actions(300)).cuda(), auto_reset=True, extras=True)
torch.cuda.synchronize()
b.check()
print("device face ok", b.last_step_kernel)
