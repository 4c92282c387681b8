"""GPU diagnostic: steps whose cycles have K = 0 (action 0,0,0) -- only the fixed per-step work runs
(state load, nozzle IK, cycle plan, reward, observation, stores).  Run under ncu to see where it goes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grasp_lab_salp_b200 import SalpBatch, default_params

n = int(os.environ.get("DIAG_N", "4096"))
b = SalpBatch(n, default_params(), seed=0)
b.reset_device()
z = torch.zeros((n, 3), device="cuda")
pl = {None: None, "0": False, "1": True}[os.environ.get("DIAG_PIPELINE")]
for _ in range(12):
    b.step_device(z, pipeline=pl)
torch.cuda.synchronize()
print("ok")
