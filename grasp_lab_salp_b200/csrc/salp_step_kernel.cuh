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
// Observation rows are assembled in a shared-memory tile ([2][32][D] floats: obs, terminal obs) and
// written out by the whole warp afterwards: consecutive lanes write consecutive floats, so an
// unsorted warp stores its 32 rows as full 128-byte lines, and the kernel never reads io.obs /
// io.terminal_obs back -- which lets salp_step_host hand it MAPPED HOST pointers (the results then
// cross PCIe as each warp finishes, overlapped with the warps still integrating).
template <int PREC>
__global__ void __launch_bounds__(32, 1)
salp_step_kernel_lat(const __grid_constant__ SalpParams p, const __grid_constant__ SalpDerived dv,
                     const __grid_constant__ SalpView v, const __grid_constant__ SalpStepIO io, uint32_t flags,
                     const int32_t* __restrict__ order) {
  extern __shared__ float tile[];
  const int D = SALP_OBS_BASE + 2 * p.num_obstacles;
  const int lane = threadIdx.x;
  const int64_t tid = (int64_t)blockIdx.x * 32 + lane;
  const bool valid = tid < v.n;
  const int64_t i = valid ? (order ? (int64_t)order[tid] : tid) : 0;
  float* obs_row = tile + lane * D;
  float* tobs_row = io.terminal_obs ? tile + (32 + lane) * D : nullptr;
  if (valid) env_step<PREC>(p, dv, v, io, flags, i, obs_row, tobs_row);
  __syncwarp();
  const int rows = __popc(__ballot_sync(0xffffffffu, valid));    // valid lanes are the low lanes
  for (int j = lane; j < 32 * D; j += 32) {                      // warp-uniform trip count (D iterations)
    const int r = j / D, k = j - r * D;
    const int64_t e = __shfl_sync(0xffffffffu, i, r);
    if (r < rows) {
      io.obs[e * D + k] = tile[j];
      if (tobs_row) io.terminal_obs[e * D + k] = tile[32 * D + j];
    }
  }
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
static inline size_t lat_tile_bytes(const SalpParams& p) { return sizeof(float) * 2 * 32 * (SALP_OBS_BASE + 2 * p.num_obstacles); }
static inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }


void salp_launch_step_f64(const SalpParams& p, const SalpView& v, const SalpStepIO& io, uint32_t flags,
                          const int32_t* order, cudaStream_t stream);
