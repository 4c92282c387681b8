// salp_common.cuh -- state layout and launch plumbing shared by the kernels and the C ABI.
//
// HBM layout (DESIGN.md "Data layout"): one RECORD per env and per scalar type,
//   f64: double [N][SALP_NUM_F64_FIELDS]   f32: float [N][NUM_F32]   i32: int32 [N][NUM_I32]
// (SALP_STATE_AOS 1).  The step kernel visits envs in K-sorted order, i.e. through a permutation:
// with a column-per-scalar layout every 8-byte access of a lane fetched its own 32-byte sector
// (measured: 6.3 KB of DRAM traffic per env-step for 1.2 KB of state, and 17 % of the kernel's time
// in long-scoreboard stalls of its prologue/epilogue); with one record per env every fetched
// sector is fully used and a lane's consecutive fields hit L1.  SALP_STATE_AOS 0 selects the
// column layout [field][N] (coalesced for an unsorted visit order) for comparison.
// The env-facing I/O arrays (actions [N,3], obs [N,D], ...) keep the row-major layout SB3 hands
// over; they are ~100 B per env-step against ~3.5e5 flop, i.e. irrelevant to the roofline.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/salp_b200.h"

#ifndef SALP_STATE_AOS
#define SALP_STATE_AOS 1
#endif
#define SALP_MAX_SUBSTEPS 4096          // cycles longer than this raise SALP_ERR_RANGE (Box actions: K <= 1348)
#define SALP_SORT_SHAPE_BINS 64          // K-sort key = (K >> 5) * 64 + min(end of shape motion >> 3, 63)
#define SALP_SORT_BINS (((SALP_MAX_SUBSTEPS >> 5) + 1) * SALP_SORT_SHAPE_BINS)
#define SALP_NUM_F32 (SALP_F32_END - SALP_F32_BASE)
#define SALP_NUM_I32 (SALP_I32_END - SALP_I32_BASE)

// Every per-env routine is `__host__ __device__` so that tests/emu/ can compile the very same
// step body for the host and check its logic against the oracle in the GPU-less CI container.
// The product library (libsalp_b200.so) only ever launches the __global__ kernels.
#define SALP_HD __host__ __device__ __forceinline__

struct SalpView {
  double* f64;          // [n][SALP_NUM_F64_FIELDS]  (records; [field][n] if !SALP_STATE_AOS)
  float* f32;           // [n][SALP_NUM_F32]
  int32_t* i32;         // [n][SALP_NUM_I32]
  int64_t n;            // envs on this GPU
  int64_t env_id_offset;
  uint64_t seed;
  // scene pool (nullable): targets [n,P,2], obstacles [n,P,nobs,2]
  const float* pool_targets;
  const float* pool_obstacles;
  int64_t pool_P;
  int32_t* status;      // sticky device status word (0 = ok)
  const double* time_table;   // t_k, k = 0..SALP_MAX_SUBSTEPS: k-fold repeated `+= dt` (robot.py:674)
  int32_t sm_count;           // SMs of the handle's device (kernel selection)
};

// Per-step scratch owned by the handle (K-sort path).
struct SalpScratch {
  int32_t* K;           // [n]   sort key of the pending cycle (K bucket, end of shape motion)
  int32_t* order;       // [n]   env indices sorted by K (descending)
  int32_t* hist;        // [SALP_SORT_BINS] counting-sort histogram / offsets
};

// Launchers implemented in salp_kernels.cu (all asynchronous on `stream`).
// Each returns the number of kernels it launched, or a negative SalpStatus.
int salp_launch_reset(const SalpParams& p, const SalpView& v, const uint8_t* mask, float* obs,
                      cudaStream_t stream);
// `kernel_name` (nullable) receives the name of the step kernel that was launched.
int salp_launch_step(const SalpParams& p, const SalpView& v, const SalpStepIO& io, uint32_t flags,
                     const SalpScratch& scratch, cudaStream_t stream, const char** kernel_name = nullptr);
int salp_launch_init(const SalpParams& p, const SalpView& v, cudaStream_t stream);
int salp_launch_trace(const SalpParams& p, const SalpView& v, int64_t env, const float action[3], double* trace,
                      int capacity, int32_t* K_out, cudaStream_t stream);
int salp_launch_ffma_probe(float* scratch, int blocks, int iters, cudaStream_t stream);

// ---- rounding-exact scalar ops -------------------------------------------------------------
// nvcc contracts a*b+c into an FMA in device code and the host compiler may keep x87/AVX
// intermediates; wherever the reference's result depends on a *separately rounded* float32 or
// float64 operation (SURVEY.md hard parts 1-2) the code says so with these.
namespace rn {
SALP_HD float fmul(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fmul_rn(a, b);
#else
  volatile float r = a * b; return r;
#endif
}
SALP_HD float fadd(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fadd_rn(a, b);
#else
  volatile float r = a + b; return r;
#endif
}
SALP_HD float fsub(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fsub_rn(a, b);
#else
  volatile float r = a - b; return r;
#endif
}
SALP_HD float fdiv(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fdiv_rn(a, b);
#else
  volatile float r = a / b; return r;
#endif
}
SALP_HD float fsqrt(float a) {
#ifdef __CUDA_ARCH__
  return __fsqrt_rn(a);
#else
  volatile float r = sqrtf(a); return r;
#endif
}
SALP_HD float ffma(float a, float b, float c) {
#ifdef __CUDA_ARCH__
  return __fmaf_rn(a, b, c);
#else
  return fmaf(a, b, c);
#endif
}
SALP_HD double dmul(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dmul_rn(a, b);
#else
  volatile double r = a * b; return r;
#endif
}
SALP_HD double dadd(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dadd_rn(a, b);
#else
  volatile double r = a + b; return r;
#endif
}
SALP_HD double dsub(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dsub_rn(a, b);
#else
  volatile double r = a - b; return r;
#endif
}
SALP_HD double ddiv(double a, double b) {
#ifdef __CUDA_ARCH__
  return __ddiv_rn(a, b);
#else
  volatile double r = a / b; return r;
#endif
}
}  // namespace rn
