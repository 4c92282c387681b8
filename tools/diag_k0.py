"""GPU diagnostic: the fixed per-step cost -- K = 0 steps (action 0,0,0) at several batch sizes, the
reset kernel as a launch-overhead yardstick; argv[1] = 'ncu' runs only a few K = 0 steps (for ncu -k)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from grasp_lab_salp_b200 import SalpBatch, default_params
dev = torch.device("cuda", 0)
def timed(fn, reps=40):
    for _ in range(5): fn()
    st = [torch.cuda.Event(enable_timing=True) for _ in range(reps)]; en = [torch.cuda.Event(enable_timing=True) for _ in range(reps)]
    for i in range(reps):
        st[i].record(); fn(); en[i].record()
    torch.cuda.synchronize()
    return float(np.median([s.elapsed_time(e) for s, e in zip(st, en)])) * 1e3
if len(sys.argv) > 1 and sys.argv[1] == "ncu":
    n = 4096
    b = SalpBatch(n, default_params(), seed=0); b.reset_device()
    z = torch.zeros((n, 3), device=dev)
    for _ in range(8): b.step_device(z)
    torch.cuda.synchronize()
    sys.exit(0)
for n in (32, 1024, 4096, 16384):
    b = SalpBatch(n, default_params(), seed=0); b.reset_device()
    z = torch.zeros((n, 3), device=dev)
    t_reset = timed(lambda: b.reset_device())
    t_def = timed(lambda: b.step_device(z)); k = b.last_step_kernel
    t_fused = timed(lambda: b.step_device(z, pipeline=False))
    t_noreset = timed(lambda: b.step_device(z, auto_reset=False, pipeline=False))
    print(f"n={n:6d}: reset kernel {t_reset:5.1f} us | K=0 step: default [{k}] {t_def:5.1f} us, fused {t_fused:5.1f} us, fused without auto-reset {t_noreset:5.1f} us")
    b.close()
