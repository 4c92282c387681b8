"""GPU diagnostic for compute-sanitizer (memcheck / racecheck): a few steps through every kernel
path at small sizes -- pipeline kernel (ragged), fused kernel natural and K-sorted, F64, randomised,
host transports."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grasp_lab_salp_b200 import PRECISION_F64, SalpBatch, default_params

rng = np.random.default_rng(0)
steps = int(os.environ.get("DIAG_STEPS", "2"))
for label, n, kw, params in [
    ("pipeline", 200, dict(pipeline=True), default_params()),
    ("fused", 200, dict(pipeline=False), default_params()),
    ("fused sorted", 200, dict(pipeline=False, sort_by_k=True), default_params()),
    ("f64", 96, dict(), default_params(precision=PRECISION_F64)),
    ("randomised", 96, dict(), default_params(randomization=31)),
]:
    b = SalpBatch(n, params, seed=1)
    b.reset()
    for t in range(steps):
        a = rng.uniform([0, 0, -1], [1, 1, 1], size=(n, 3)).astype(np.float32)
        b.step(a, auto_reset=True, **kw)
    b.check()
    print(label, "ok", float(b.reward.mean()))
    b.close()
