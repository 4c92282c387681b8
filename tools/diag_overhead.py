"""GPU diagnostic: fixed per-step cost of the step kernel (prologue: state load, nozzle IK, cycle
plan; epilogue: reward, observation, auto-reset) = time of a step whose cycle has K = 0 substeps
(action (0, 0, yaw)), next to short uniform-K steps."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grasp_lab_salp_b200 import SalpBatch, default_params

PIPE = {None: None, '0': False, '1': True}[os.environ.get('DIAG_PIPELINE')]
GENERIC = os.environ.get('DIAG_GENERIC') == '1'
dev = torch.device("cuda", 0)
n = int(os.environ.get("DIAG_N", "4096"))
b = SalpBatch(n, default_params(), seed=0)
b.reset_device()
g = torch.Generator(device=dev)
g.manual_seed(1)


def timed(actions, label, steps=50):
    for _ in range(10):
        b.step_device(actions, pipeline=PIPE, generic=GENERIC)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for i in range(steps):
        b.step_device(actions, pipeline=PIPE, generic=GENERIC)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(steps))[steps // 2]
    k = b.dev["substeps"].float().mean().item() if "substeps" in b.dev else -1
    print(f"{label:40s} median {ms * 1e3:8.1f} us")


z = torch.zeros((n, 3), device=dev)
timed(z, "K=0 (action 0,0,0)")
a = z.clone(); a[:, 2] = torch.rand(n, generator=g, device=dev) * 2 - 1
timed(a, "turn-only cycles (0,0,yaw~U)")
a = z.clone(); a[:, 0] = 0.5
timed(a, "a0=0.5, no coast")
a = z.clone(); a[:, 0] = 0.5; a[:, 1] = 0.5
timed(a, "a0=0.5, coast 5 s (uniform K)")
a = z.clone(); a[:, 0] = 1.0; a[:, 1] = 1.0
timed(a, "a0=1, coast 10 s (K=1348 everywhere)")
