// salp_step_kernel.cuh -- the step kernel template, shared by salp_kernels.cu (MIXED) and
// salp_step_f64.cu (F64; that translation unit is compiled with -fmad=false so that the
// reference-mode arithmetic is not contracted into FMAs the reference does not perform).
#pragma once
#include "salp_env.cuh"

// Two register budgets of the same kernel: the throughput build (<= 128 registers, 16 warps/SM) for
// batches that fill the GPU, and the latency build (<= 255 registers, used with one warp per
// block) for small batches, where each SM sub-partition holds at most one warp and only
// instruction-level parallelism inside that warp hides the FP32 pipe latency.
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

// Kernel choice by batch size.  Up to SALP_LAT_MAX_ENVS the latency build (one warp per block,
// <= 255 registers, 9 warps per SM) is used; beyond that the throughput build (128-thread blocks,
// 128 registers, 16 warps per SM).  The crossover was measured (tools/diag_crossover.py); the
// environment variable SALP_LAT_MAX_ENVS overrides it for experiments.
#include <cstdlib>
static inline int64_t salp_lat_max_envs() {
  static int64_t cached = -1;
  if (cached < 0) {
    const char* e = getenv("SALP_LAT_MAX_ENVS");
    cached = e ? atoll(e) : (int64_t)148 * 4 * 32;
  }
  return cached;
}
static inline int block_for(int64_t n) { return n <= salp_lat_max_envs() ? 32 : 128; }
static inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }


void salp_launch_step_f64(const SalpParams& p, const SalpView& v, const SalpStepIO& io, uint32_t flags,
                          const int32_t* order, cudaStream_t stream);
