"""GPU diagnostic: axisymmetric vs general form of the loop, per-column differences after one step."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grasp_lab_salp_b200 import SalpBatch, default_params
from grasp_lab_salp_b200.params import FIELDS

n = 512
rng = np.random.default_rng(3)
for label, acts in (("uniform", rng.uniform([0, 0, -1], [1, 1, 1], size=(n, 3)).astype(np.float32)),
                    ("coast only", np.tile(np.array([[0.0, 0.5, 0.3]], np.float32), (n, 1))),
                    ("contraction, no coast", np.tile(np.array([[0.5, 0.0, 0.3]], np.float32), (n, 1)))):
    for pipeline in (True, False):
        a, b = SalpBatch(n, default_params(), seed=2), SalpBatch(n, default_params(), seed=2)
        a.reset(); b.reset()
        a.step(acts, pipeline=pipeline)
        b.step(acts, pipeline=pipeline, generic=True)
        bad = []
        for col in FIELDS:
            x, y = a.get_state(col).astype(np.float64), b.get_state(col).astype(np.float64)
            d = np.flatnonzero(~((x == y) | (np.isnan(x) & np.isnan(y))))
            if d.size:
                bad.append((col, d.size, float(np.abs(x - y)[d].max())))
        print(label, "pipeline" if pipeline else "fused", bad[:12])
