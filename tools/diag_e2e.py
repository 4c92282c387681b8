"""GPU diagnostic: wall-clock ms per SalpBatch.step (host buffers through salp_step_host) at several
batch sizes.  SALP_ZERO_COPY=0 forces the staged transport."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grasp_lab_salp_b200 import SalpBatch, default_params
from grasp_lab_salp_b200.params import sort_by_k_auto

for n in [int(x) for x in os.environ.get("DIAG_NS", "4096,16384,65536,262144,1048576").split(",")]:
    b = SalpBatch(n, default_params(), seed=0)
    b.reset()
    rng = np.random.default_rng(1)
    acts = [b.host_buffer((n, 3), np.float32) for _ in range(4)]
    for a in acts:
        a[:] = rng.uniform([0, 0, -1], [1, 1, 1], size=(n, 3))
    sort = sort_by_k_auto(n)
    for i in range(int(os.environ.get('DIAG_WARMUP', '300')) if n <= 65536 else 5):
        b.step(acts[i % 4], auto_reset=True, sort_by_k=sort, extras=False)
    steps = 60 if n <= 65536 else 15
    t0 = time.perf_counter()
    for i in range(steps):
        b.step(acts[i % 4], auto_reset=True, sort_by_k=sort, extras=False)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    print(f"n={n:8d}: e2e {dt * 1e3:.4f} ms/step  {n / dt / 1e6:.2f} M env-steps/s  (zero_copy={os.environ.get('SALP_ZERO_COPY', '1')})")
    b.close()
