"""Host side of the tensor-core LSTM cell (csrc/salp_lstm.cu): the rollout-side forward of one
`torch.nn.LSTMCell(obs_dim, 256)` of the MlpLstmPolicy the reference trains
(src/train_robot_recurrent_ppo.py:100-105), with sb3_contrib's episode-start state reset folded in.

The learner keeps the fp32 torch cell (autograd, BPTT); rollouts call `LstmCellB200.step`, which
updates the per-env state in place.  bf16 operands, fp32 accumulation and fp32 cell update: the
rollout-side h differs from the fp32 cell by ~1e-3 (tests/test_ppo.py).  CUDA only -- there is no
CPU path.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from . import _lib

HIDDEN = 256


class LstmCellB200:
    def __init__(self, cell: nn.LSTMCell, n_envs: int):
        if cell.hidden_size != HIDDEN:
            raise ValueError(f"the tensor-core cell is built for lstm_hidden_size = {HIDDEN}")
        dev = cell.weight_hh.device
        if dev.type != "cuda":
            raise RuntimeError("LstmCellB200 needs the cell's parameters on a CUDA device (no CPU path)")
        self.lib = _lib.load()
        self.cell, self.device, self.n = cell, dev, int(n_envs)
        self.obs_dim = cell.input_size
        self.packed = torch.empty(self.lib.salp_lstm_weight_bytes() // 2, dtype=torch.bfloat16, device=dev)
        self.bias = torch.empty(4 * HIDDEN, dtype=torch.float32, device=dev)
        self.scratch = torch.empty(self.lib.salp_lstm_scratch_bytes(self.n) // 2, dtype=torch.bfloat16, device=dev)
        self.pack()

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def pack(self):
        """Re-read the torch cell's parameters (call after every optimiser update, before the rollout)."""
        c = self.cell
        rc = self.lib.salp_lstm_pack_weights(C.c_void_p(c.weight_ih.data_ptr()), C.c_void_p(c.weight_hh.data_ptr()),
                                             C.c_void_p(c.bias_ih.data_ptr()), C.c_void_p(c.bias_hh.data_ptr()),
                                             self.obs_dim, HIDDEN, C.c_void_p(self.packed.data_ptr()),
                                             C.c_void_p(self.bias.data_ptr()), self._stream())
        if rc != 0:
            raise RuntimeError(f"salp_lstm_pack_weights failed ({rc})")

    def step(self, obs, starts, h, c, h_out=None, c_out=None):
        """(h, c) <- cell(obs, (h * keep, c * keep)), keep = ~starts; in place unless h_out / c_out are given.
        obs [N, D] fp32, starts [N] bool or None, h / c [N, 256] fp32, all contiguous on the cell's device."""
        h_out = h if h_out is None else h_out
        c_out = c if c_out is None else c_out
        n = obs.shape[0]
        if not (0 < n <= self.n and obs.shape[1] == self.obs_dim and obs.dtype == torch.float32 and obs.is_contiguous()):
            raise ValueError(f"obs must be a contiguous float32 [n <= {self.n}, {self.obs_dim}] tensor")
        for t in (h, c, h_out, c_out):
            if not (tuple(t.shape) == (n, HIDDEN) and t.dtype == torch.float32 and t.is_contiguous() and t.is_cuda):
                raise ValueError(f"h / c must be contiguous float32 [{n}, {HIDDEN}] CUDA tensors")
        if starts is not None and not (tuple(starts.shape) == (n,) and starts.element_size() == 1 and starts.is_contiguous()):
            raise ValueError("starts must be a contiguous bool / uint8 [n] tensor")
        rc = self.lib.salp_lstm_cell(C.c_void_p(self.packed.data_ptr()), C.c_void_p(self.bias.data_ptr()),
                                     C.c_void_p(obs.data_ptr()), C.c_void_p(starts.data_ptr()) if starts is not None else None,
                                     C.c_void_p(h.data_ptr()), C.c_void_p(c.data_ptr()), C.c_void_p(h_out.data_ptr()),
                                     C.c_void_p(c_out.data_ptr()), C.c_void_p(self.scratch.data_ptr()), n, self.obs_dim, HIDDEN,
                                     self._stream())
        if rc != 0:
            raise RuntimeError(f"salp_lstm_cell failed ({rc})")
        return h_out, c_out

    def check(self):
        if self.lib.salp_lstm_check() != 0:
            raise RuntimeError("salp_lstm_cell: a kernel gave up waiting on one of its barriers")
