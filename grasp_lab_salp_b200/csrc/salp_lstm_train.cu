// salp_lstm_train.cu -- the element-wise halves of the LEARNER's LSTM step (RecurrentPPO update, BPTT
// over the rollout; src/train_robot_recurrent_ppo.py:85-107 configures sb3_contrib's MlpLstmPolicy).
//
// The learner keeps fp32 operands and cuBLAS GEMMs (gradients want them), but a step of
// torch.nn.LSTMCell under autograd is ~25 launches forward and ~50 backward, and the update replays
// 32 steps x 2 cells x 160 minibatch steps per iteration: launch-bound (1.68 s per iteration for
// 8192 envs against 0.015 s for the rollout).  grasp_lab_salp_b200/lstm_seq.py restates the sequence
// as an autograd.Function with TWO launches per cell and step in each direction -- one GEMM and one of
// the kernels below -- and batches everything that does not depend on the recurrence (input projection,
// heads, weight gradients) over the whole [T x B] block.
//
//   forward :  gates [B, 4H] (= x W_ih^T + b + (h keep) W_hh^T, torch gate order i, f, g, o)
//              -> act = (sigma(i), sigma(f), tanh(g), sigma(o)), c' = f (c keep) + i g, h' = o tanh(c'),
//                 and h' keep_next: the masked operand of the NEXT step's GEMM (episode-start reset of
//                 sb3_contrib's _process_sequence)
//   backward:  dh = dh_ext + dh_rec keep_next, dc' from the later step
//              -> dgates [B, 4H], dc keep (gradient w.r.t. the previous cell state)
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/salp_b200.h"

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(256)
salp_lstm_pointwise_fwd_kernel(const float* __restrict__ gates, const float* __restrict__ c_prev,
                               const float* __restrict__ keep_cur, const float* __restrict__ keep_next, int64_t total, int H,
                               float* __restrict__ act, float* __restrict__ c_out, float* __restrict__ h_out,
                               float* __restrict__ hm_next) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int64_t b = idx / H;
  const int u = (int)(idx - b * H);
  const float* g = gates + b * 4 * H;
  const float i_ = sigmoid_acc(g[u]), f_ = sigmoid_acc(g[H + u]), g_ = tanhf(g[2 * H + u]), o_ = sigmoid_acc(g[3 * H + u]);
  const float cm = c_prev[idx] * (keep_cur ? keep_cur[b] : 1.f);
  const float c = fmaf(f_, cm, i_ * g_);
  const float h = o_ * tanhf(c);
  float* a = act + b * 4 * H;
  a[u] = i_; a[H + u] = f_; a[2 * H + u] = g_; a[3 * H + u] = o_;
  c_out[idx] = c;
  h_out[idx] = h;
  if (hm_next) hm_next[idx] = h * (keep_next ? keep_next[b] : 1.f);
}

__global__ void __launch_bounds__(256)
salp_lstm_pointwise_bwd_kernel(const float* __restrict__ dh_ext, const float* __restrict__ dh_rec,
                               const float* __restrict__ keep_next, const float* __restrict__ dc_next,
                               const float* __restrict__ act, const float* __restrict__ c_cur,
                               const float* __restrict__ c_prev, const float* __restrict__ keep_cur, int64_t total, int H,
                               float* __restrict__ dgates, float* __restrict__ dc_prev) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int64_t b = idx / H;
  const int u = (int)(idx - b * H);
  const float* a = act + b * 4 * H;
  const float i_ = a[u], f_ = a[H + u], g_ = a[2 * H + u], o_ = a[3 * H + u];
  float dh = dh_ext ? dh_ext[idx] : 0.f;
  if (dh_rec) dh = fmaf(dh_rec[idx], keep_next ? keep_next[b] : 1.f, dh);
  const float tc = tanhf(c_cur[idx]);
  const float kc = keep_cur ? keep_cur[b] : 1.f;
  const float cm = c_prev[idx] * kc;
  float dc = dh * o_ * (1.f - tc * tc);
  if (dc_next) dc += dc_next[idx];
  float* d = dgates + b * 4 * H;
  d[u] = dc * g_ * i_ * (1.f - i_);
  d[H + u] = dc * cm * f_ * (1.f - f_);
  d[2 * H + u] = dc * i_ * (1.f - g_ * g_);
  d[3 * H + u] = dh * tc * o_ * (1.f - o_);
  dc_prev[idx] = dc * f_ * kc;
}

extern "C" {

int salp_lstm_pointwise_fwd(const float* gates_dev, const float* c_prev_dev, const float* keep_cur_dev,
                            const float* keep_next_dev, int64_t batch, int32_t hidden, float* act_dev, float* c_out_dev,
                            float* h_out_dev, float* hm_next_dev, void* stream) {
  if (!gates_dev || !c_prev_dev || !act_dev || !c_out_dev || !h_out_dev || batch <= 0 || hidden <= 0) return SALP_ERR_INVALID;
  const int64_t total = batch * hidden;
  salp_lstm_pointwise_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      gates_dev, c_prev_dev, keep_cur_dev, keep_next_dev, total, hidden, act_dev, c_out_dev, h_out_dev, hm_next_dev);
  return cudaPeekAtLastError() == cudaSuccess ? SALP_OK : SALP_ERR_CUDA;
}

int salp_lstm_pointwise_bwd(const float* dh_ext_dev, const float* dh_rec_dev, const float* keep_next_dev,
                            const float* dc_next_dev, const float* act_dev, const float* c_cur_dev, const float* c_prev_dev,
                            const float* keep_cur_dev, int64_t batch, int32_t hidden, float* dgates_dev, float* dc_prev_dev,
                            void* stream) {
  if (!act_dev || !c_cur_dev || !c_prev_dev || !dgates_dev || !dc_prev_dev || batch <= 0 || hidden <= 0) return SALP_ERR_INVALID;
  const int64_t total = batch * hidden;
  salp_lstm_pointwise_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      dh_ext_dev, dh_rec_dev, keep_next_dev, dc_next_dev, act_dev, c_cur_dev, c_prev_dev, keep_cur_dev, total, hidden,
      dgates_dev, dc_prev_dev);
  return cudaPeekAtLastError() == cudaSuccess ? SALP_OK : SALP_ERR_CUDA;
}

}  // extern "C"
