// Micro-benchmark (measurement tooling, not product): cycles per substep of the pieces of the MIXED
// substep loop for ONE warp per SM sub-partition -- the regime of BASELINE's 4096-env config, where
// the step time is K_max x (cycles one warp needs per substep).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/_bin/ubench_substep tools/ubench_substep.cu
#include <cstdio>
#include <cstring>
#include "../grasp_lab_salp_b200/csrc/salp_env.cuh"

static SalpParams defaults() {
  SalpParams p; memset(&p, 0, sizeof p);
  const double pi = 3.14159265358979323846;
  p.nozzle_length1 = p.nozzle_length2 = p.nozzle_length3 = 0.05; p.nozzle_area = 0.00016; p.nozzle_mass = 1.0;
  p.nozzle_gamma = pi / 4; p.nozzle_angle_speed = 31 * pi / 30;
  p.dry_mass = 1.0; p.init_length = 0.3; p.init_width = 0.15; p.max_contraction = 0.06; p.density = 1000.0; p.dt = 0.01;
  p.buoy_mass = 0.195; p.skin_mass = 0.145; p.tube_mass = 0.414; p.tube_volume = pi * ((0.058 / 2) * (0.058 / 2)) * 0.15;
  p.discharge_coefficient = 0.3; p.drag_force_ratio = 0.25; p.drag_torque_ratio = 0.1;
  const double amf[3] = {0.5, 0.6, 0.6}, amrf[3] = {0.2, 0.2, 0.2}, amt[3] = {0.3, 0.6, 0.6};
  const double tdr[6] = {1.5, 2.5, 2.5, 1.5, 2.5, 1.5}, rdr[6] = {0.1, 0.3, 0.5, 0.2, 0.5, 0.2};
  for (int i = 0; i < 3; i++) { p.added_mass_force[i] = amf[i]; p.added_mass_rate_force[i] = amrf[i]; p.added_mass_torque[i] = amt[i]; p.added_mass_rate_torque[i] = amrf[i]; }
  for (int i = 0; i < 6; i++) { p.trans_drag_range[i] = tdr[i]; p.rot_drag_range[i] = rdr[i]; }
  p.num_obstacles = 2; p.precision = SALP_PRECISION_MIXED;
  return p;
}

__device__ __forceinline__ void init_state(const SalpDerived& dv, Coef32& g, Motion32& s) {
  const float dir[3] = {-1.f, 0.01f * threadIdx.x, 0.002f};
  make_coefs<true>(dv, dir, false, 0.15f, 0.075f, 0.f, 0.f, 0.f, -0.0193f, 0.f, 0.f, g);
  s.v0 = 0.3f + 0.001f * threadIdx.x; s.v1 = -0.05f; s.v2 = 1e-4f; s.w0 = 1e-3f; s.w1 = -2e-3f; s.w2 = 0.3f;
  s.ac0 = s.ac1 = s.ac2 = s.al0 = s.al1 = s.al2 = 0.f;
  s.phi_lo = s.theta_lo = s.psi_lo = 0.f; s.sph = 0.01f; s.cph = 0.99995f; s.sth = -0.02f; s.cth = 0.9998f; s.sps = 0.6f; s.cps = 0.8f;
  s.pw0 = s.pw1 = s.pw2 = s.pos0 = s.pos1 = s.pos2 = s.ang0 = s.ang1 = s.ang2 = 0.f; s.vw0 = s.vw1 = 0.f;
}
__device__ __forceinline__ float fold(const Motion32& s) {
  return s.v0 + s.v1 + s.v2 + s.w0 + s.w1 + s.w2 + s.pw0 + s.pw1 + s.pw2 + s.pos0 + s.pos1 + s.pos2 + s.ang0 + s.ang1 + s.ang2 +
         s.sph + s.cph + s.sth + s.cth + s.sps + s.cps + s.phi_lo + s.theta_lo + s.psi_lo;
}

// MODE 0: dyn only   1: kin_world only   2: kin_step + dyn (fused coast loop)   3: dyn + kin_body
// 4: dyn + kin_body + STS of (v, w)      5: LDS of (v, w) + kin_world
template <int MODE, int UNROLL>
__global__ void __launch_bounds__(128, 1) bench(const __grid_constant__ SalpDerived dv, float* out, long long* cyc, int iters) {
  __shared__ float4 ra[4][16][32];
  __shared__ float2 rb[4][16][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Coef32 g; Motion32 s;
  init_state(dv, g, s);
  for (int k = 0; k < 16; k++) { ra[warp][k][lane] = make_float4(s.v0, s.v1, s.v2, s.w0); rb[warp][k][lane] = make_float2(s.w1, s.w2); }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it += UNROLL) {
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      const int j = it + u;
      if (MODE == 0) dyn_step<false, true, true>(dv, g, s);
      if (MODE == 1) kin_world(dv, s);
      if (MODE == 2) { kin_step(dv, s); dyn_step<false, true, true>(dv, g, s); }
      if (MODE == 3) { kin_body(dv, s); dyn_step<false, true, true>(dv, g, s); }
      if (MODE == 4) {
        kin_body(dv, s); dyn_step<false, true, true>(dv, g, s);
        ra[warp][j & 15][lane] = make_float4(s.v0, s.v1, s.v2, s.w0); rb[warp][j & 15][lane] = make_float2(s.w1, s.w2);
      }
      if (MODE == 5) {
        const float4 a = ra[warp][j & 15][lane]; const float2 b = rb[warp][j & 15][lane];
        kin_world(dv, s);
        s.v0 = a.x; s.v1 = a.y; s.v2 = a.z; s.w0 = a.w; s.w1 = b.x; s.w2 = b.y;
      }
    }
  }
  long long t1 = clock64();
  float f = fold(s);
  if (f == 12345.678f) out[0] = f;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

// The dyn -> kin hand-off of the four-warp pipeline kernel in isolation: warp 0 produces (v, w) per
// substep into a ring, warp 1 consumes; chunks of C substeps, NBUF chunks in flight, named barriers.
__device__ __forceinline__ void bar_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }
template <int C, int NBUF>
__global__ void __launch_bounds__(64, 1) bench_pair(const __grid_constant__ SalpDerived dv, float* out, long long* cyc, int iters) {
  __shared__ float4 ra[C * NBUF][32];
  __shared__ float2 rb[C * NBUF][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Coef32 g; Motion32 s;
  init_state(dv, g, s);
  __syncthreads();
  long long t0 = clock64();
  const int nch = iters / C;
  if (warp == 0) {
    for (int c = 0; c < nch; c++) {
      if (c >= NBUF) bar_sync(1 + NBUF + c % NBUF);
#pragma unroll
      for (int u = 0; u < C; u++) {
        const int j = c * C + u;
        kin_body(dv, s); dyn_step<false, true, true>(dv, g, s);
        ra[j % (C * NBUF)][lane] = make_float4(s.v0, s.v1, s.v2, s.w0); rb[j % (C * NBUF)][lane] = make_float2(s.w1, s.w2);
      }
      __syncwarp();
      bar_arrive(1 + c % NBUF);
    }
  } else {
    for (int c = 0; c < nch; c++) {
      bar_sync(1 + c % NBUF);
#pragma unroll
      for (int u = 0; u < C; u++) {
        const int j = c * C + u;
        const float4 a = ra[j % (C * NBUF)][lane]; const float2 b = rb[j % (C * NBUF)][lane];
        kin_world(dv, s);
        s.v0 = a.x; s.v1 = a.y; s.v2 = a.z; s.w0 = a.w; s.w1 = b.x; s.w2 = b.y;
      }
      __syncwarp();
      bar_arrive(1 + NBUF + c % NBUF);
    }
  }
  long long t1 = clock64();
  __syncthreads();
  float f = fold(s);
  if (f == 12345.678f) out[0] = f;
  if (lane == 0 && blockIdx.x == 0) cyc[warp] = t1 - t0;
}
template <int C, int NBUF>
void run_pair(const SalpDerived& dv, float* out, long long* cyc) {
  const int iters = 4096;
  for (int rep = 0; rep < 2; rep++) { bench_pair<C, NBUF><<<148, 64>>>(dv, out, cyc, iters); cudaDeviceSynchronize(); }
  long long c[2]; cudaMemcpy(c, cyc, 16, cudaMemcpyDeviceToHost);
  printf("dyn -> kin pair, chunk %2d x %d buffers: dyn warp %7.1f, kin warp %7.1f cycles per substep\n", C, NBUF,
         (double)c[0] / iters, (double)c[1] / iters);
}

// producer stages of the pipeline kernels in isolation (shape moving all the time: a0 = 1, refill ramp)
#include "../grasp_lab_salp_b200/csrc/salp_pipe_kernel.cuh"
template <int STAGE, int PER_TRIP>
__global__ void __launch_bounds__(32, 1) bench_producer(const __grid_constant__ SalpParams p, const __grid_constant__ SalpDerived dv,
                                                         const double* __restrict__ table, float* out, long long* cyc, int iters) {
  __shared__ float ring1[8][32][8];
  __shared__ float ring2[8][32][20];
  const int lane = threadIdx.x;
  CyclePlan plan = make_cycle_plan(p, 1.0f, 0.0f, 0.3f + 0.001f * lane, 0.0, 0.0);
  PhasePlan pp = make_phase_plan(plan, table, dv.inv_dt);
  Body64 b; memset(&b, 0, sizeof b);
  b.length = p.init_length; b.width = p.init_width; b.phase = 3;
  b.prev_volume = ellipsoid_volume(b.length, b.width) - p.tube_volume;
  const float dir[3] = {(float)plan.dir[0], (float)plan.dir[1], (float)plan.dir[2]};
  ShapeTrack st; Coef32 g;
  mixed_init_shape<true>(p, dv, b, dir, st, g);
  ShapeFront f; f.dl = 0.01f; f.I_rate0 = 1e-4f; f.I_rate1 = 2e-4f; f.dV_dt = 1e-3f; f.com = -0.019f; f.com_rate = 1e-3f; f.com_acc = 1e-2f; f.jet_on = 0.f;
  for (int k = 0; k < 8; k++) front_store(f, &ring1[k][lane][0]);
  __syncwarp();
  double tj = table[1];
  float acc = 0.f;
  long long t0 = clock64();
  for (int j = 1; j <= iters; j += PER_TRIP) {
    if (STAGE == 0) {        // front: fp64 shape chain + differences -> ring 1
#pragma unroll
      for (int u = 0; u < PER_TRIP; u++) {
        ShapeFront ff;
        shape_front(p, dv, plan, tj, ((j + u) & 255) + 1, pp.k_T0, pp.k_jet, st, ff);
        tj = rn::dadd(tj, p.dt); if (((j + u) & 255) == 255) tj = table[1];
        front_store(ff, &ring1[(j + u) & 7][lane][0]);
      }
    } else {                 // coefs: ring 1 -> fp32 coefficient set -> ring 2
#pragma unroll
      for (int u = 0; u < PER_TRIP; u++) {
        ShapeFront ff; Coef32 gg;
        front_load(ff, &ring1[(j + u) & 7][lane][0]);
        make_coefs<true>(dv, dir, ff, gg);
        coef_store<true>(gg, &ring2[(j + u) & 7][lane][0]);
      }
    }
  }
  long long t1 = clock64();
  acc += (float)st.s.V + ring2[3][lane][5] + ring1[2][lane][1];
  if (acc == 12345.678f) out[0] = acc;
  if (lane == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int STAGE, int PER_TRIP>
void run_producer(const char* label, const SalpParams& p, const SalpDerived& dv, const double* table, float* out, long long* cyc) {
  const int iters = 4096;
  for (int rep = 0; rep < 2; rep++) { bench_producer<STAGE, PER_TRIP><<<148, 32>>>(p, dv, table, out, cyc, iters); cudaDeviceSynchronize(); }
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-20s %d update(s) per trip: %7.1f cycles per substep\n", label, PER_TRIP, (double)c / iters);
}

// the same pair through the pipeline kernel's own chunk runner (segments, flushes every 32 substeps)
#include "../grasp_lab_salp_b200/csrc/salp_pipe4_kernel.cuh"
template <int VAR>
__global__ void __launch_bounds__(64, 1) bench_pair_runner(const __grid_constant__ SalpDerived dv, float* out, long long* cyc, int Kall) {
  constexpr int C = SALP_P4_CHUNK, NBUF = SALP_P4_NBUF3, SLOTS = C * NBUF;
  __shared__ float4 ra[SLOTS][32];
  __shared__ float2 rb[SLOTS][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Coef32 g; Motion32 s; Body64 b; memset(&b, 0, sizeof b);
  init_state(dv, g, s);
  const int K = Kall, Kw = Kall;
  const int nch3 = (Kw - 1 + C - 1) / C;
  __syncthreads();
  long long t0 = clock64();
  if (warp == 0) {
    auto it = [&](int j, bool at32) {
      kin_body(dv, s); dyn_step<false, true, true>(dv, g, s);
      ra[j % SLOTS][lane] = make_float4(s.v0, s.v1, s.v2, s.w0); rb[j % SLOTS][lane] = make_float2(s.w1, s.w2);
      if (VAR != 1 && at32) flush_body(b, s);
    };
    int dn = 0;
    for (int c = 0; c < nch3; c++) {
      if (c >= NBUF) bar_sync(1 + NBUF + c % NBUF);
      const int j0 = c * C + 1, je = (c + 1) * C < Kw - 1 ? (c + 1) * C : Kw - 1;
      if (VAR == 2) {
        if (je - j0 + 1 == C) {
#pragma unroll
          for (int u = 0; u < C; u++) it(j0 + u, u == C - 1 && (je & 31) == 0);
        } else for (int j = j0; j <= je; j++) it(j, (j & 31) == 0);
      } else p4_run_chunk(j0, je, K, dn, it);
      __syncwarp();
      bar_arrive(1 + c % NBUF);
    }
  } else {
    auto it = [&](int j, bool at32) {
      const float4 a = ra[j % SLOTS][lane]; const float2 bb = rb[j % SLOTS][lane];
      kin_world(dv, s);
      if (VAR != 1 && at32) flush_world(b, s);
      s.v0 = a.x; s.v1 = a.y; s.v2 = a.z; s.w0 = a.w; s.w1 = bb.x; s.w2 = bb.y;
    };
    int dn = 0;
    for (int c = 0; c < nch3; c++) {
      bar_sync(1 + c % NBUF);
      const int j0 = c * C + 1, je = (c + 1) * C < Kw - 1 ? (c + 1) * C : Kw - 1;
      if (VAR == 2) {
        if (je - j0 + 1 == C) {
#pragma unroll
          for (int u = 0; u < C; u++) it(j0 + u, u == C - 1 && (je & 31) == 0);
        } else for (int j = j0; j <= je; j++) it(j, (j & 31) == 0);
      } else p4_run_chunk(j0, je, K, dn, it);
      __syncwarp();
      bar_arrive(1 + NBUF + c % NBUF);
    }
  }
  long long t1 = clock64();
  __syncthreads();
  float f = fold(s) + (float)(b.pw[0] + b.pos[0] + b.eul[0]);
  if (f == 12345.678f) out[0] = f;
  if (lane == 0 && blockIdx.x == 0) cyc[warp] = t1 - t0;
}

template <int MODE, int UNROLL>
void run(const char* label, const SalpDerived& dv, int warps, float* out, long long* cyc) {
  const int iters = 4096;
  for (int rep = 0; rep < 2; rep++) { bench<MODE, UNROLL><<<148, 32 * warps>>>(dv, out, cyc, iters); cudaDeviceSynchronize(); }
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-34s unroll %d, %d warp(s)/SM: %7.1f cycles per substep\n", label, UNROLL, warps, (double)c / iters);
}

int main() {
  SalpParams p = defaults();
  SalpDerived dv = make_derived(p);
  float* out; long long* cyc; cudaMalloc(&out, 4); cudaMalloc(&cyc, 64);
  {
    double h[SALP_MAX_SUBSTEPS + 1]; volatile double acc = 0.0;
    for (int k = 0; k <= SALP_MAX_SUBSTEPS; k++) { h[k] = acc; acc = acc + p.dt; }
    double* table; cudaMalloc(&table, sizeof h); cudaMemcpy(table, h, sizeof h, cudaMemcpyHostToDevice);
    run_producer<0, 1>("front (fp64 chain)", p, dv, table, out, cyc); run_producer<0, 2>("front (fp64 chain)", p, dv, table, out, cyc);
    run_producer<0, 4>("front (fp64 chain)", p, dv, table, out, cyc);
    run_producer<1, 1>("coefs (fp32 set)", p, dv, table, out, cyc); run_producer<1, 2>("coefs (fp32 set)", p, dv, table, out, cyc);
    run_producer<1, 4>("coefs (fp32 set)", p, dv, table, out, cyc);
  }
#define RUNNER(VAR, label) do { for (int rep = 0; rep < 2; rep++) { bench_pair_runner<VAR><<<148, 64>>>(dv, out, cyc, 4096); cudaDeviceSynchronize(); } \
    long long c[2]; cudaMemcpy(c, cyc, 16, cudaMemcpyDeviceToHost); \
    printf("pair, %-46s (chunk %d): dyn %7.1f, kin %7.1f cycles per substep\n", label, SALP_P4_CHUNK, c[0] / 4096.0, c[1] / 4096.0); } while (0)
  RUNNER(0, "p4_run_chunk + flushes");
  RUNNER(1, "p4_run_chunk, no flushes");
  RUNNER(2, "plain unrolled chunk + flushes");
  run_pair<8, 2>(dv, out, cyc); run_pair<8, 3>(dv, out, cyc); run_pair<8, 4>(dv, out, cyc); run_pair<16, 2>(dv, out, cyc);
  run_pair<4, 4>(dv, out, cyc); run_pair<32, 2>(dv, out, cyc);
  run<0, 1>("dyn", dv, 1, out, cyc); run<0, 4>("dyn", dv, 1, out, cyc); run<0, 8>("dyn", dv, 1, out, cyc);
  run<1, 1>("kin_world", dv, 1, out, cyc); run<1, 4>("kin_world", dv, 1, out, cyc); run<1, 8>("kin_world", dv, 1, out, cyc);
  run<2, 1>("kin_step + dyn (fused)", dv, 1, out, cyc); run<2, 2>("kin_step + dyn (fused)", dv, 1, out, cyc); run<2, 4>("kin_step + dyn (fused)", dv, 1, out, cyc);
  run<3, 1>("dyn + kin_body", dv, 1, out, cyc); run<3, 8>("dyn + kin_body", dv, 1, out, cyc);
  run<4, 1>("dyn + kin_body + STS", dv, 1, out, cyc); run<4, 8>("dyn + kin_body + STS", dv, 1, out, cyc);
  run<5, 1>("LDS + kin_world", dv, 1, out, cyc); run<5, 8>("LDS + kin_world", dv, 1, out, cyc);
  run<2, 1>("kin_step + dyn (fused)", dv, 4, out, cyc); run<0, 8>("dyn", dv, 4, out, cyc); run<1, 8>("kin_world", dv, 4, out, cyc);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
