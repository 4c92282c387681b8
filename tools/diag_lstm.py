"""GPU: the tcgen05 LSTM cell (csrc/salp_lstm.cu) against torch.nn.LSTMCell -- error and time.

  python tools/diag_lstm.py [N ...]        one JSON line per batch size
"""
import json
import sys

import torch
from torch import nn

from grasp_lab_salp_b200.lstm import LstmCellB200

H, D = 256, 10


def reference(cell, obs, starts, h, c, bf16_operands):
    """float64 restatement of nn.LSTMCell with the episode-start reset; optionally with the operands
    rounded to bf16 exactly as the kernel rounds them."""
    keep = (~starts).double().unsqueeze(-1)
    r = (lambda t: t.float().bfloat16().double()) if bf16_operands else (lambda t: t.double())
    hk = (h.double() * keep).float()
    g = r(obs) @ r(cell.weight_ih).T + r(hk) @ r(cell.weight_hh).T + (cell.bias_ih.float() + cell.bias_hh.float()).double()
    i, f, gg, o = g.chunk(4, dim=1)
    c2 = torch.sigmoid(f) * (c.double() * keep) + torch.sigmoid(i) * torch.tanh(gg)
    return torch.sigmoid(o) * torch.tanh(c2), c2


def timeit(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3      # microseconds


def main():
    sizes = [int(x) for x in sys.argv[1:]] or [100, 8192, 65536]
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    cell = nn.LSTMCell(D, H).to(dev)
    with torch.no_grad():
        for p in cell.parameters():
            p.mul_(1.5)
    for n in sizes:
        obs = torch.randn(n, D, device=dev)
        h = torch.tanh(torch.randn(n, H, device=dev))
        c = torch.randn(n, H, device=dev)
        starts = torch.rand(n, device=dev) < 0.1
        fused = LstmCellB200(cell, n)
        h1, c1 = torch.empty_like(h), torch.empty_like(c)
        fused.step(obs, starts, h, c, h1, c1)
        fused.check()
        row = {"n": n}
        with torch.no_grad():
            for name, flag in (("bf16_operands", True), ("fp32", False)):
                hr, cr = reference(cell, obs, starts, h, c, flag)
                row[f"max_err_h_vs_{name}"] = float((h1.double() - hr).abs().max())
                row[f"max_err_c_vs_{name}"] = float((c1.double() - cr).abs().max())
            keep = (~starts).float().unsqueeze(-1)
            ht, ct = cell(obs, (h * keep, c * keep))
            row["max_err_h_vs_torch_cell"] = float((h1 - ht).abs().max())
            # in place
            h2, c2 = h.clone(), c.clone()
            fused.step(obs, starts, h2, c2)
            row["in_place_identical"] = bool(torch.equal(h2, h1) and torch.equal(c2, c1))
            row["us_fused"] = timeit(lambda: fused.step(obs, starts, h, c, h1, c1))
            row["us_torch_fp32"] = timeit(lambda: cell(obs, (h * keep, c * keep)))
            torch.backends.cuda.matmul.allow_tf32 = True
            row["us_torch_tf32"] = timeit(lambda: cell(obs, (h * keep, c * keep)))
            torch.backends.cuda.matmul.allow_tf32 = False
            with torch.autocast("cuda", dtype=torch.bfloat16):
                row["us_torch_bf16_autocast"] = timeit(lambda: cell(obs, (h * keep, c * keep)))
        fused.check()
        flop = 2.0 * n * (H + D) * 4 * H
        row["tflops_fused"] = flop / row["us_fused"] * 1e-6
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
