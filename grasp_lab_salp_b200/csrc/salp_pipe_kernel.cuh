// salp_pipe_kernel.cuh -- the small-batch step kernel: a warp-specialised, feed-forward pipeline.
//
// With a few thousand envs the GPU is almost empty (4096 envs = 128 warps on 592 SM sub-partitions)
// and the step time is K_max (~1340 substeps of the slowest env) x the latency of ONE warp's
// substep.  The fused kernel issues ~250 instructions per substep from a single warp.  But the
// substep is feed-forward:
//
//     shape(j)  ->  dyn(j)  ->  kin(j)
//
//   * the body shape and every coefficient derived from it depend on the action and on j only,
//     never on the motion state;
//   * the Newton/Euler equations + velocity update (dyn) need the coefficients and (v, w);
//   * the Euler angles / world position / body-frame integrals (kin) only consume (v, w) and never
//     feed back (there is no gravity or current in the reference's model).
//
// So one block of three warps owns 32 envs: warp 2 produces coefficient sets ahead of time, warp 0
// runs the ~85-instruction dyn recurrence (the true critical path), warp 1 integrates the
// kinematics behind it.  The stages talk through two shared-memory rings, indexed by substep,
// with chunk-granular (16 substeps) double-buffered hand-off on named barriers
// (bar.arrive on the producer side, bar.sync on the consumer side), so the critical warp never
// waits unless a producer has fallen a full chunk behind.  Each warp sits on its own SM
// sub-partition (warp id % 4).  Results equal the fused kernel's up to the grouping of the fp32
// chunk sums (tests/test_gpu_parity.py::test_pipeline_kernel_matches_fused_kernel).
#pragma once
#include "salp_env.cuh"

#define SALP_PIPE_CHUNK 16
#define SALP_PIPE_SLOTS (2 * SALP_PIPE_CHUNK)
#define SALP_PIPE_NCOEF 26
#define SALP_PIPE_THREADS 96

struct PipeShared {
  float ringA[SALP_PIPE_SLOTS][SALP_PIPE_NCOEF][32];   // shape -> dyn : Coef32 of substep j, slot j % SLOTS
  float ringB[SALP_PIPE_SLOTS][6][32];                 // dyn -> kin   : (v, w) after substep j
  double merge[22][32];                                // kin / shape results for the epilogue (warp 0)
};

__device__ __forceinline__ void pipe_bar_sync(int id) {
  asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
}
__device__ __forceinline__ void pipe_bar_arrive(int id) {
  asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory");
}
// named barriers (0 is __syncthreads): full/empty x double buffer, for both rings
#define PIPE_FULL_A(b) (1 + (b))
#define PIPE_EMPTY_A(b) (3 + (b))
#define PIPE_FULL_B(b) (5 + (b))
#define PIPE_EMPTY_B(b) (7 + (b))

__device__ __forceinline__ void coef_store(const Coef32& g, float (*slot)[32], int lane) {
  int f = 0;
#pragma unroll
  for (int i = 0; i < 3; i++) slot[f++][lane] = g.aj[i];
#pragma unroll
  for (int i = 0; i < 3; i++) slot[f++][lane] = g.kdm[i];
#pragma unroll
  for (int i = 0; i < 3; i++) slot[f++][lane] = g.mrm[i];
  slot[f++][lane] = g.com; slot[f++][lane] = g.com_rate; slot[f++][lane] = g.com_acc;
  slot[f++][lane] = g.tj1; slot[f++][lane] = g.tj2;
#pragma unroll
  for (int i = 0; i < 3; i++) slot[f++][lane] = g.kqI[i];
#pragma unroll
  for (int i = 0; i < 3; i++) slot[f++][lane] = g.klI[i];
#pragma unroll
  for (int i = 0; i < 3; i++) slot[f++][lane] = g.JdI[i];
#pragma unroll
  for (int i = 0; i < 3; i++) slot[f++][lane] = g.AdI[i];
}
__device__ __forceinline__ void coef_load(Coef32& g, const float (*slot)[32], int lane) {
  int f = 0;
#pragma unroll
  for (int i = 0; i < 3; i++) g.aj[i] = slot[f++][lane];
#pragma unroll
  for (int i = 0; i < 3; i++) g.kdm[i] = slot[f++][lane];
#pragma unroll
  for (int i = 0; i < 3; i++) g.mrm[i] = slot[f++][lane];
  g.com = slot[f++][lane]; g.com_rate = slot[f++][lane]; g.com_acc = slot[f++][lane];
  g.tj1 = slot[f++][lane]; g.tj2 = slot[f++][lane];
#pragma unroll
  for (int i = 0; i < 3; i++) g.kqI[i] = slot[f++][lane];
#pragma unroll
  for (int i = 0; i < 3; i++) g.klI[i] = slot[f++][lane];
#pragma unroll
  for (int i = 0; i < 3; i++) g.JdI[i] = slot[f++][lane];
#pragma unroll
  for (int i = 0; i < 3; i++) g.AdI[i] = slot[f++][lane];
}

__global__ void __launch_bounds__(SALP_PIPE_THREADS, 1)
salp_step_kernel_pipe(const __grid_constant__ SalpParams p, const __grid_constant__ SalpDerived dv,
                      const __grid_constant__ SalpView v, const __grid_constant__ SalpStepIO io, uint32_t flags) {
  extern __shared__ __align__(16) unsigned char pipe_smem[];
  PipeShared& sh = *reinterpret_cast<PipeShared*>(pipe_smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * 32 + lane;
  const bool live = i < v.n;
  constexpr int C = SALP_PIPE_CHUNK;

  // Every warp reads the env's action and state itself (reads only; all writes happen in warp 0's
  // epilogue after the block-wide barrier) and derives the same integer plan.
  StepCtx cx;
  Body64 b;
  int Kraw = 0;
  PhasePlan pp;
  pp.k_ref = pp.k_T0 = pp.k_jet = pp.upd_a_end = pp.upd_b_begin = pp.upd_b_end = 0;
  if (live) {
    env_step_begin(p, v, io, i, cx, b);
    Kraw = plan_substeps(cx.plan, v.time_table);
    if (Kraw > 0) pp = make_phase_plan(cx.plan, v.time_table, dv.inv_dt);
  }
  const int K = Kraw > 0 ? Kraw : 0;
  const int Kw = __reduce_max_sync(0xffffffffu, K);
  const int nchunks = Kw / C + 1;                          // chunks of update/substep indices j = 0..Kw

  if (warp == 2) {
    // ---------------- shape warp: coefficient sets, ahead of the dyn warp ----------------
    ShapeTrack st;
    Coef32 g;
    const float dir[3] = {(float)cx.plan.dir[0], (float)cx.plan.dir[1], (float)cx.plan.dir[2]};
    int next_upd = 0x7fffffff;
    if (K > 0) {
      mixed_init_shape(p, dv, b, dir, st, g);
      next_upd = 0;
    }
    for (int c = 0; c < nchunks; c++) {
      const int bsel = c & 1;
      if (c >= 2) pipe_bar_sync(PIPE_EMPTY_A(bsel));
      const int jend = (c * C + C - 1 < Kw) ? c * C + C - 1 : Kw;
      for (int j = c * C; j <= jend; j++) {
        if (j == next_upd && j <= K) {
          if (j > 0) shape_update(p, dv, cx.plan, v.time_table, dir, j, pp.k_T0, pp.k_jet, st, g);
          if (j < K) coef_store(g, sh.ringA[j % SALP_PIPE_SLOTS], lane);
          next_upd = j == 0 ? 1 : next_update_after(j, pp);
        }
      }
      pipe_bar_arrive(PIPE_FULL_A(bsel));
    }
    if (K > 0) mixed_finish_shape(p, st, K, b);
    sh.merge[13][lane] = b.length; sh.merge[14][lane] = b.width; sh.merge[15][lane] = b.prev_volume;
    sh.merge[16][lane] = b.prevI[0]; sh.merge[17][lane] = b.prevI[1];
    sh.merge[18][lane] = b.com; sh.merge[19][lane] = b.com_rate; sh.merge[20][lane] = b.prev_com_rate;
    sh.merge[21][lane] = b.com_acc;
  } else if (warp == 0) {
    // ---------------- dyn warp: the critical recurrence ----------------
    Motion32 s;
    Coef32 g;
    mixed_init_dyn(b, s);
    int next_upd = 0;
    for (int c = 0; c < nchunks; c++) {
      const int bsel = c & 1;
      pipe_bar_sync(PIPE_FULL_A(bsel));
      if (c >= 2) pipe_bar_sync(PIPE_EMPTY_B(bsel));
      const int jend = (c * C + C - 1 < Kw) ? c * C + C - 1 : Kw;
      for (int j = c * C; j <= jend; j++) {
        if (j < K) {
          if (j == next_upd) {
            coef_load(g, sh.ringA[j % SALP_PIPE_SLOTS], lane);
            next_upd = j == 0 ? 1 : next_update_after(j, pp);
          }
          dyn_step(dv, g, s);
          float(*slot)[32] = sh.ringB[j % SALP_PIPE_SLOTS];
          slot[0][lane] = s.v0; slot[1][lane] = s.v1; slot[2][lane] = s.v2;
          slot[3][lane] = s.w0; slot[4][lane] = s.w1; slot[5][lane] = s.w2;
        }
      }
      pipe_bar_arrive(PIPE_EMPTY_A(bsel));
      pipe_bar_arrive(PIPE_FULL_B(bsel));
    }
    if (K > 0) mixed_finish_dyn(s, b);
  } else {
    // ---------------- kin warp: Euler angles, world position, body-frame integrals ----------------
    Motion32 s;
    mixed_init_kin(b, s);
    for (int c = 0; c < nchunks; c++) {
      const int bsel = c & 1;
      pipe_bar_sync(PIPE_FULL_B(bsel));
      const int jend = (c * C + C - 1 < Kw) ? c * C + C - 1 : Kw;
      for (int j = c * C; j <= jend; j++) {
        if (j < K) {
          const float(*slot)[32] = sh.ringB[j % SALP_PIPE_SLOTS];
          s.v0 = slot[0][lane]; s.v1 = slot[1][lane]; s.v2 = slot[2][lane];
          s.w0 = slot[3][lane]; s.w1 = slot[4][lane]; s.w2 = slot[5][lane];
          kin_step(dv, s);
        }
      }
      if (c * C < K) flush_chunk(b, s);
      pipe_bar_arrive(PIPE_EMPTY_B(bsel));
    }
    if (K > 0) b.speed_world = (double)sqrtf(s.vw0 * s.vw0 + s.vw1 * s.vw1);
#pragma unroll
    for (int k = 0; k < 3; k++) {
      sh.merge[k][lane] = b.pw[k]; sh.merge[3 + k][lane] = b.pos[k];
      sh.merge[6 + k][lane] = b.ang[k]; sh.merge[9 + k][lane] = b.eul[k];
    }
    sh.merge[12][lane] = b.speed_world;
  }
  __syncthreads();
  if (warp == 0 && live) {
    const double pos0[3] = {b.pos[0], b.pos[1], b.pos[2]};
    const double ang0[3] = {b.ang[0], b.ang[1], b.ang[2]};
#pragma unroll
    for (int k = 0; k < 3; k++) {
      b.pw[k] = sh.merge[k][lane]; b.pos[k] = sh.merge[3 + k][lane];
      b.ang[k] = sh.merge[6 + k][lane]; b.eul[k] = sh.merge[9 + k][lane];
    }
    b.speed_world = sh.merge[12][lane];
    b.length = sh.merge[13][lane]; b.width = sh.merge[14][lane]; b.prev_volume = sh.merge[15][lane];
    b.prevI[0] = sh.merge[16][lane]; b.prevI[1] = sh.merge[17][lane]; b.prevI[2] = K > 0 ? sh.merge[17][lane] : b.prevI[2];
    b.com = sh.merge[18][lane]; b.com_rate = sh.merge[19][lane]; b.prev_com_rate = sh.merge[20][lane];
    b.com_acc = sh.merge[21][lane];
    double t = 0.0;
    if (K > 0) {
      b.prev_com = b.com;
      t = v.time_table[K];
      b.phase = phase_at(cx.plan, t);
    }
    env_step_end(p, v, io, flags, i, cx, pos0, ang0, b, Kraw, t);
  }
}
