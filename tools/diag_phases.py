"""GPU diagnostic: step time of uniform batches that isolate the fixed cost (K = 0), the shape-moving
substeps (a0 = 1 without coast: K = 348) and the coast substeps (a0 = 1 with 10 s of coast: K = 1348),
for the default kernel and the fused one, plus uniform-random actions (L2 not flushed)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from grasp_lab_salp_b200 import SalpBatch, default_params
dev = torch.device("cuda", 0)
n = int(os.environ.get("DIAG_N", "4096"))
def run(acts, **kw):
    b = SalpBatch(n, default_params(), seed=0); b.reset_device()
    for i in range(5): b.step_device(acts[i % len(acts)], **kw)
    st = [torch.cuda.Event(enable_timing=True) for _ in range(40)]; en = [torch.cuda.Event(enable_timing=True) for _ in range(40)]
    for i in range(40):
        st[i].record(); b.step_device(acts[i % len(acts)], **kw); en[i].record()
    torch.cuda.synchronize(); b.check()
    ms = float(np.median([s.elapsed_time(e) for s, e in zip(st, en)])); k = b.last_step_kernel; b.close()
    return ms * 1e3, k
z = torch.zeros((1, n, 3), device=dev)
a348 = z.clone(); a348[..., 0] = 1.0
a1348 = a348.clone(); a1348[..., 1] = 1.0
g = torch.Generator(device=dev); g.manual_seed(1234)
u = torch.rand((16, n, 3), generator=g, device=dev); u[..., 2] = u[..., 2] * 2 - 1
only = os.environ.get("DIAG_ONLY")
for label, kw in (("default", {}), ("fused", dict(pipeline=False))):
    if only and label != only: continue
    t0, k = run(z, **kw); tA, _ = run(a348, **kw); tB, _ = run(a1348, **kw); tu, _ = run(u, **kw)
    print(f"{label:8s} [{k}] n={n}: K=0 {t0:6.1f} us | moving {1965*(tA-t0)/348:6.1f} cyc/substep | coast {1965*(tB-tA)/1000:6.1f} cyc/substep | "
          f"K=1348 {tB:6.1f} us | uniform random {tu:6.1f} us", flush=True)
