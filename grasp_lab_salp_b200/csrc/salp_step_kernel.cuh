// salp_step_kernel.cuh -- the step kernel template, shared by salp_kernels.cu (MIXED) and
// salp_step_f64.cu (F64; that translation unit is compiled with -fmad=false so that the
// reference-mode arithmetic is not contracted into FMAs the reference does not perform).
#pragma once
#include "salp_env.cuh"

// Two register budgets of the same kernel.  The wide build (one warp per block, <= 255 registers,
// 8 warps per SM) is the default for MIXED at every batch size: the substep loop is bound by
// instruction issue, not by latency, so 8 warps with the whole working set in registers run as
// fast as 16 warps at 128 registers on full GPUs (1M envs: 6.48 vs 6.51 ms) and up to 1.45x
// faster on partial waves (42k envs: 0.37 vs 0.53 ms; profiles/README.md, "kernel choice").
// The 128-register build remains for F64 (FP64-pipe bound) and as an experiment switch.
template <int PREC>
__global__ void __launch_bounds__(32, 1)
salp_step_kernel_lat(const __grid_constant__ SalpParams p, const __grid_constant__ SalpDerived dv,
                     const __grid_constant__ SalpView v, const __grid_constant__ SalpStepIO io, uint32_t flags,
                     const int32_t* __restrict__ order) {
  int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= v.n) return;
  int64_t i = order ? (int64_t)order[tid] : tid;
  env_step<PREC>(p, dv, v, io, flags, i);
}

template <int PREC>
__global__ void __launch_bounds__(128, 4)
salp_step_kernel(const __grid_constant__ SalpParams p, const __grid_constant__ SalpDerived dv,
                 const __grid_constant__ SalpView v, const __grid_constant__ SalpStepIO io, uint32_t flags,
                 const int32_t* __restrict__ order) {
  int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= v.n) return;
  int64_t i = order ? (int64_t)order[tid] : tid;
  env_step<PREC>(p, dv, v, io, flags, i);
}

// Kernel choice: the wide build up to SALP_LAT_MAX_ENVS envs (default: always); the environment
// variable exists to reproduce the crossover measurement of tools/diag_crossover.py.
#include <cstdlib>
static inline int64_t salp_lat_max_envs() {
  static int64_t cached = -1;
  if (cached < 0) {
    const char* e = getenv("SALP_LAT_MAX_ENVS");
    cached = e ? atoll(e) : INT64_MAX;
  }
  return cached;
}
static inline int block_for(int64_t n) { return n <= salp_lat_max_envs() ? 32 : 128; }
static inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }


void salp_launch_step_f64(const SalpParams& p, const SalpView& v, const SalpStepIO& io, uint32_t flags,
                          const int32_t* order, cudaStream_t stream);
