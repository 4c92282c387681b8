"""The learner's LSTM over a whole rollout as ONE autograd.Function (RecurrentPPO update, BPTT).

`torch.nn.LSTMCell` stepped T times under autograd is ~25 launches per step forward and ~50
backward; the update of BASELINE config 4 (8192 envs, T = 32, two cells, 160 minibatch steps per
iteration) is launch-bound: 1.68 s per iteration against 0.015 s for the rollout.  Here a cell costs
TWO launches per step in each direction -- one cuBLAS GEMM on the recurrence and one hand-written
element-wise kernel (csrc/salp_lstm_train.cu) -- and everything that does not depend on the recurrence
is batched over the [T x B] block: the input projection in front, the weight / bias gradients behind.
Same fp32 arithmetic as the cell loop (sums in another order: 1e-6), same semantics: sb3_contrib's
`_process_sequence` -- the state is reset where an episode starts (`keep = 1 - episode_start`).

On CPU tensors (the GPU-less test container, gloo runs) the element-wise halves are plain torch ops.
"""
from __future__ import annotations

import ctypes as C

import torch


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _pointwise_fwd(gates, c_prev, keep_cur, keep_next, act, c_out, h_out, hm_next):
    if gates.is_cuda:
        from . import _lib
        B, H = c_prev.shape
        rc = _lib.load().salp_lstm_pointwise_fwd(_ptr(gates), _ptr(c_prev), _ptr(keep_cur), _ptr(keep_next), B, H, _ptr(act),
                                                 _ptr(c_out), _ptr(h_out), _ptr(hm_next),
                                                 C.c_void_p(torch.cuda.current_stream(gates.device).cuda_stream))
        if rc != 0:
            raise RuntimeError(f"salp_lstm_pointwise_fwd failed ({rc})")
        return
    i, f, g, o = gates.chunk(4, dim=1)
    i, f, g, o = torch.sigmoid(i), torch.sigmoid(f), torch.tanh(g), torch.sigmoid(o)
    c = f * (c_prev * keep_cur.unsqueeze(1)) + i * g
    h = o * torch.tanh(c)
    act.copy_(torch.cat([i, f, g, o], dim=1))
    c_out.copy_(c)
    h_out.copy_(h)
    if hm_next is not None:
        hm_next.copy_(h * keep_next.unsqueeze(1))


def _pointwise_bwd(dh_ext, dh_rec, keep_next, dc_next, act, c_cur, c_prev, keep_cur, dgates, dc_prev):
    if act.is_cuda:
        from . import _lib
        B, H = c_prev.shape
        rc = _lib.load().salp_lstm_pointwise_bwd(_ptr(dh_ext), _ptr(dh_rec), _ptr(keep_next), _ptr(dc_next), _ptr(act),
                                                 _ptr(c_cur), _ptr(c_prev), _ptr(keep_cur), B, H, _ptr(dgates), _ptr(dc_prev),
                                                 C.c_void_p(torch.cuda.current_stream(act.device).cuda_stream))
        if rc != 0:
            raise RuntimeError(f"salp_lstm_pointwise_bwd failed ({rc})")
        return
    i, f, g, o = act.chunk(4, dim=1)
    dh = dh_ext if dh_rec is None else dh_ext + dh_rec * keep_next.unsqueeze(1)
    tc = torch.tanh(c_cur)
    kc = keep_cur.unsqueeze(1)
    dc = dh * o * (1 - tc * tc)
    if dc_next is not None:
        dc = dc + dc_next
    dgates.copy_(torch.cat([dc * g * i * (1 - i), dc * (c_prev * kc) * f * (1 - f), dc * i * (1 - g * g),
                            dh * tc * o * (1 - o)], dim=1))
    dc_prev.copy_(dc * f * kc)


class LstmSequence(torch.autograd.Function):
    """hs [T, B, H] = the hidden states of nn.LSTMCell(D, H) run over x [T, B, D] from (h0, c0), with the
    state multiplied by keep[t] [B] before step t.  Gradients: w_ih, w_hh, b_ih, b_hh (x, h0, c0 and
    keep are data)."""

    @staticmethod
    def forward(ctx, x, h0, c0, keep, w_ih, w_hh, b_ih, b_hh):
        T, B, D = x.shape
        H = h0.shape[1]
        x = x.contiguous()
        keep = keep.contiguous()
        ig = torch.addmm(b_ih + b_hh, x.reshape(T * B, D), w_ih.t()).view(T, B, 4 * H)
        act = x.new_empty((T, B, 4 * H))
        cs = x.new_empty((T + 1, B, H))
        hs = x.new_empty((T, B, H))
        hm = x.new_empty((T, B, H))          # h_{t-1} keep[t]: operand of step t's GEMM
        cs[0].copy_(c0)
        torch.mul(h0, keep[0].unsqueeze(1), out=hm[0])
        for t in range(T):
            gates = torch.addmm(ig[t], hm[t], w_hh.t())
            last = t + 1 == T
            _pointwise_fwd(gates, cs[t], keep[t], None if last else keep[t + 1], act[t], cs[t + 1], hs[t],
                           None if last else hm[t + 1])
        ctx.save_for_backward(x, keep, w_ih, w_hh, act, cs, hm)
        return hs

    @staticmethod
    def backward(ctx, dhs):
        x, keep, w_ih, w_hh, act, cs, hm = ctx.saved_tensors
        T, B, D = x.shape
        H = hm.shape[2]
        dhs = dhs.contiguous()
        dg = x.new_empty((T, B, 4 * H))
        dc = [x.new_empty((B, H)), x.new_empty((B, H))]
        dh_rec = None
        for t in reversed(range(T)):
            last = t + 1 == T
            _pointwise_bwd(dhs[t], dh_rec, None if last else keep[t + 1], None if last else dc[(t + 1) & 1], act[t], cs[t + 1],
                           cs[t], keep[t], dg[t], dc[t & 1])
            if t > 0:
                dh_rec = torch.mm(dg[t], w_hh)          # gradient w.r.t. hm[t] = h_{t-1} keep[t]
        dg2 = dg.view(T * B, 4 * H)
        dw_hh = torch.mm(dg2.t(), hm.view(T * B, H))
        dw_ih = torch.mm(dg2.t(), x.reshape(T * B, D))
        db = dg2.sum(0)
        return None, None, None, None, dw_ih, dw_hh, db, db.clone()


def lstm_sequence(cell: torch.nn.LSTMCell, x, h0, c0, keep):
    return LstmSequence.apply(x, h0, c0, keep, cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh)
