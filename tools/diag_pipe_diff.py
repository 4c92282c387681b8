"""GPU diagnostic: pipeline kernel vs fused kernel, per-column worst difference."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from grasp_lab_salp_b200 import SalpBatch, default_params
from grasp_lab_salp_b200.params import FIELDS

n, T = 1000, 4
rng = np.random.default_rng(12)
acts = rng.uniform([0, 0, -1], [1, 1, 1], size=(T, n, 3)).astype(np.float32)
pipe = SalpBatch(n, default_params(), seed=2)
fused = SalpBatch(n, default_params(), seed=2)
pipe.reset(); fused.reset()
for t in range(T):
    o1, r1, te1, tr1 = pipe.step(acts[t], auto_reset=False, pipeline=True)
    o2, r2, te2, tr2 = fused.step(acts[t], auto_reset=False, pipeline=False)
    print(f"step {t}: substeps equal {np.array_equal(pipe.substeps, fused.substeps)}  obs max diff {np.abs(o1 - o2).max():.3e}  "
          f"reward {np.abs(r1 - r2).max():.3e}")
    for col in FIELDS:
        a, b = pipe.get_state(col).astype(np.float64), fused.get_state(col).astype(np.float64)
        bad = np.flatnonzero(a != b)
        if bad.size:
            e = np.abs(a - b)[bad]
            print(f"   {col:18s} {bad.size:5d} envs differ, max abs {e.max():.3e}, rel {(e / np.maximum(np.abs(b[bad]), 1e-30)).max():.3e}, first env {bad[0]} K={fused.substeps[bad[0]]}")
    if t == 0:
        break
