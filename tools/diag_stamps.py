"""GPU diagnostic: where block 0 of the small-batch pipeline kernel spends one step (clock64 stamps of
its dyn warp).  Cases: K = 0 (fixed cost only), K = 1348, uniform-random actions (L2 flushed before the
stamped step, like bench.py)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from grasp_lab_salp_b200 import SalpBatch, _lib, default_params

dev = torch.device("cuda", 0)
n = int(os.environ.get("DIAG_N", "4096"))
lib = _lib.load()
NAMES = ["entry->begin", "begin->plan", "plan->init", "init->loops_done", "loops_done->sync", "sync->step_end", "step_end->obs_out"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(label, acts, do_flush):
    b = SalpBatch(n, default_params(), seed=0)
    b.reset_device()
    rows = []
    for i in range(12):
        if do_flush:
            flush.zero_()
        b.step_device(acts[i % len(acts)], extra_flags=0x20000000)
        out = (C.c_longlong * 16)()
        assert lib.salp_debug_p4_stamps(out) == 0
        if i >= 4:
            rows.append([out[k + 1] - out[k] for k in range(7)] + [out[7] - out[0], out[8], out[9]])
    b.check()
    b.close()
    r = np.median(np.array(rows, dtype=np.float64), axis=0)
    parts = " | ".join(f"{nm} {c / 1965:6.2f} us" for nm, c in zip(NAMES, r[:7]))
    print(f"{label:28s} total {r[7] / 1965:6.1f} us (Kmax {int(r[8])}, W {int(r[9])}): {parts}", flush=True)


z = torch.zeros((1, n, 3), device=dev)
a1348 = z.clone(); a1348[..., 0] = 1.0; a1348[..., 1] = 1.0
g = torch.Generator(device=dev); g.manual_seed(1234)
u = torch.rand((16, n, 3), generator=g, device=dev); u[..., 2] = u[..., 2] * 2 - 1
for fl in (False, True):
    tag = "flushed" if fl else "warm"
    run(f"K=0 {tag}", z, fl)
    run(f"K=1348 {tag}", a1348, fl)
    run(f"uniform random {tag}", u, fl)
