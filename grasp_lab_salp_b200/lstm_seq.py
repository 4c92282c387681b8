"""The learner's LSTM over a whole rollout as ONE autograd.Function (RecurrentPPO update, BPTT).

`torch.nn.LSTMCell` stepped T times under autograd is ~25 launches per step forward and ~50
backward; the update of BASELINE config 4 (8192 envs, T = 32, two cells, 160 minibatch steps per
iteration) is launch-bound: 1.68 s per iteration against 0.015 s for the rollout.  Here the actor's and the critic's cell TOGETHER cost
two launches per step in each direction -- one batched cuBLAS GEMM on the recurrence and one hand-written
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
    """G independent nn.LSTMCell(D, H) run side by side over x [T, B, D] (the actor's and the critic's LSTM
    see the same observations): hs [T, G, B, H] from (h0, c0) [G, B, H], the state multiplied by
    keep[t] [B] before step t.  w_ih [G, 4H, D], w_hh [G, 4H, H], b_ih / b_hh [G, 4H].  One batched GEMM
    and one element-wise launch per step and direction for ALL G cells.  Gradients: the four weight
    arguments (x, h0, c0 and keep are data)."""

    @staticmethod
    def forward(ctx, x, h0, c0, keep, w_ih, w_hh, b_ih, b_hh):
        T, B, D = x.shape
        G, _, H = h0.shape
        x = x.contiguous()
        keep_g = keep.unsqueeze(1).expand(T, G, B).reshape(T, G * B).contiguous()       # one row per (cell, env)
        x2 = x.reshape(T * B, D)
        # input projection of every step at once: [G, T*B, 4H]
        ig = torch.baddbmm((b_ih + b_hh).unsqueeze(1), x2.unsqueeze(0).expand(G, T * B, D), w_ih.transpose(1, 2)).view(G, T, B, 4 * H)
        w_hh_t = w_hh.transpose(1, 2)
        act = x.new_empty((T, G * B, 4 * H))
        cs = x.new_empty((T + 1, G * B, H))
        hs = x.new_empty((T, G * B, H))
        hm = x.new_empty((T, G * B, H))          # h_{t-1} keep[t]: operand of step t's GEMM
        cs[0].copy_(c0.reshape(G * B, H))
        torch.mul(h0.reshape(G * B, H), keep_g[0].unsqueeze(1), out=hm[0])
        for t in range(T):
            gates = torch.baddbmm(ig[:, t], hm[t].view(G, B, H), w_hh_t)                 # [G, B, 4H], contiguous
            last = t + 1 == T
            _pointwise_fwd(gates.view(G * B, 4 * H), cs[t], keep_g[t], None if last else keep_g[t + 1], act[t], cs[t + 1], hs[t],
                           None if last else hm[t + 1])
        ctx.save_for_backward(x, keep_g, w_ih, w_hh, act, cs, hm)
        ctx.G = G
        return hs.view(T, G, B, H)

    @staticmethod
    def backward(ctx, dhs):
        x, keep_g, w_ih, w_hh, act, cs, hm = ctx.saved_tensors
        G = ctx.G
        T, B, D = x.shape
        H = hm.shape[2]
        dhs = dhs.contiguous().view(T, G * B, H)
        dg = x.new_empty((T, G * B, 4 * H))
        dc = [x.new_empty((G * B, H)), x.new_empty((G * B, H))]
        dh_rec = None
        for t in reversed(range(T)):
            last = t + 1 == T
            _pointwise_bwd(dhs[t], dh_rec, None if last else keep_g[t + 1], None if last else dc[(t + 1) & 1], act[t], cs[t + 1],
                           cs[t], keep_g[t], dg[t], dc[t & 1])
            if t > 0:
                dh_rec = torch.bmm(dg[t].view(G, B, 4 * H), w_hh).view(G * B, H)       # gradient w.r.t. hm[t] = h_{t-1} keep[t]
        dgt = dg.view(T, G, B, 4 * H).permute(1, 3, 0, 2).reshape(G, 4 * H, T * B)       # [G, 4H, T*B]
        dw_hh = torch.bmm(dgt, hm.view(T, G, B, H).permute(1, 0, 2, 3).reshape(G, T * B, H))
        dw_ih = torch.bmm(dgt, x.reshape(1, T * B, D).expand(G, T * B, D))
        db = dgt.sum(2)
        return None, None, None, None, dw_ih, dw_hh, db, db.clone()


def lstm_sequence(cells, x, h0, c0, keep):
    """cells: list of G nn.LSTMCell of equal shape; h0 / c0: lists of G tensors [B, H].  Returns the G
    hidden-state sequences [T, B, H]."""
    st = torch.stack
    hs = LstmSequence.apply(x, st(list(h0)), st(list(c0)), keep, st([c.weight_ih for c in cells]), st([c.weight_hh for c in cells]),
                            st([c.bias_ih for c in cells]), st([c.bias_hh for c in cells]))
    return [hs[:, g] for g in range(len(cells))]
