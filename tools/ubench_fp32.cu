// Micro-benchmark (measurement tooling): issue rate and latency of FFMA vs the packed FFMA2 / FMUL2 /
// FADD2 (fma.rn.f32x2, PTX ISA 8.6, sm_100+) for ONE warp per SM sub-partition and for several.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/ubench_fp32 tools/ubench_fp32.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float ffma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float mufu_sqrt(float a) { float d; asm volatile("sqrt.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(a)); return d; }

template <int MODE, int ILP>
__global__ void bench(float* out, long long* cyc, int iters, float fa, float fb) {
  float x[ILP]; u64 y[ILP];
  const float a = fa + threadIdx.x * 1e-9f, b = fb;
  u64 a2, b2; { float2 t = make_float2(a, a); a2 = *reinterpret_cast<u64*>(&t); t = make_float2(b, b); b2 = *reinterpret_cast<u64*>(&t); }
#pragma unroll
  for (int i = 0; i < ILP; i++) { x[i] = threadIdx.x + i; float2 t = make_float2(x[i], x[i] + 1); y[i] = *reinterpret_cast<u64*>(&t); }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int i = 0; i < ILP; i++) {
        if (MODE == 0) x[i] = ffma1(x[i], a, b);
        if (MODE == 1) y[i] = ffma2(y[i], a2, b2);
        if (MODE == 2) x[i] = mufu_sqrt(x[i]);
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) { float2 t = *reinterpret_cast<float2*>(&y[i]); s += x[i] + t.x + t.y; }
  if (s == 12345.678f) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int MODE, int ILP>
void run(const char* label, int warps_per_block, float* out, long long* cyc) {
  const int iters = 2000;
  bench<MODE, ILP><<<148, 32 * warps_per_block>>>(out, cyc, iters, 0.999f, 0.001f);
  cudaDeviceSynchronize();
  bench<MODE, ILP><<<148, 32 * warps_per_block>>>(out, cyc, iters, 0.999f, 0.001f);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  double per = (double)c / (iters * 8.0 * ILP);
  printf("%-10s ILP=%d warps/block=%2d : %.2f cycles per warp-instruction (per warp)  -> %.2f instr/cycle/SMSP\n", label, ILP,
         warps_per_block, per, (warps_per_block / 4.0 < 1 ? 1 : warps_per_block / 4.0) / per);
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 4); cudaMalloc(&cyc, 8);
  run<0, 1>("FFMA", 1, out, cyc); run<0, 2>("FFMA", 1, out, cyc); run<0, 4>("FFMA", 1, out, cyc); run<0, 8>("FFMA", 1, out, cyc);
  run<1, 1>("FFMA2", 1, out, cyc); run<1, 2>("FFMA2", 1, out, cyc); run<1, 4>("FFMA2", 1, out, cyc); run<1, 8>("FFMA2", 1, out, cyc);
  run<0, 8>("FFMA", 4, out, cyc); run<1, 8>("FFMA2", 4, out, cyc);
  run<0, 8>("FFMA", 8, out, cyc); run<1, 8>("FFMA2", 8, out, cyc);
  run<0, 8>("FFMA", 16, out, cyc); run<1, 8>("FFMA2", 16, out, cyc);
  run<0, 8>("FFMA", 32, out, cyc); run<1, 8>("FFMA2", 32, out, cyc);
  run<2, 1>("MUFU.SQRT", 1, out, cyc); run<2, 4>("MUFU.SQRT", 1, out, cyc);
  return 0;
}
