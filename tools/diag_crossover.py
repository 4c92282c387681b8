"""Diagnostic: sorted vs unsorted step time around the batch sizes where the kernel choice flips."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from grasp_lab_salp_b200 import SalpBatch, default_params

PIPE = {None: None, '0': False, '1': True}[os.environ.get('DIAG_PIPELINE')]
GENERIC = os.environ.get('DIAG_GENERIC') == '1'
dev = torch.device('cuda', 0)
import os
for n in [int(x) for x in os.environ.get('DIAG_NS','18944,24576,32768,42624,49152,65536,98304,131072').split(',')]:
    b = SalpBatch(n, default_params(), seed=0)
    b.reset_device()
    g = torch.Generator(device=dev)
    g.manual_seed(1234)
    u = torch.rand((8, n, 3), generator=g, device=dev)
    u[..., 2] = u[..., 2] * 2 - 1
    out = []
    for sort in (False, True):
        for i in range(4):
            b.step_device(u[i % 8], sort_by_k=sort, pipeline=PIPE, generic=GENERIC)
        steps = 30
        st = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        en = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        for i in range(steps):
            st[i].record()
            b.step_device(u[i % 8], sort_by_k=sort, pipeline=PIPE, generic=GENERIC)
            en[i].record()
        torch.cuda.synchronize()
        out.append(np.mean([s.elapsed_time(e) for s, e in zip(st, en)]))
    print(f"n={n:6d}: unsorted {out[0]:.3f} ms ({n / out[0] / 1e3:.1f} M/s)   sorted {out[1]:.3f} ms ({n / out[1] / 1e3:.1f} M/s)")
    b.close()
