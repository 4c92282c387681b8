"""GPU diagnostic for ncu: a few steps of one uniform batch.  argv: n a0 a1 [fused]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from grasp_lab_salp_b200 import SalpBatch, default_params
n, a0, a1 = int(sys.argv[1]), float(sys.argv[2]), float(sys.argv[3])
pipeline = False if len(sys.argv) > 4 and sys.argv[4] == "fused" else None
b = SalpBatch(n, default_params(), seed=0); b.reset_device()
a = torch.zeros((n, 3), device="cuda"); a[:, 0] = a0; a[:, 1] = a1
for _ in range(8): b.step_device(a, pipeline=pipeline)
torch.cuda.synchronize(); b.check(); print("ok", b.last_step_kernel)
