// salp_pipe_kernel.cuh -- the small-batch step kernel: a four-warp feed-forward pipeline.
//
// With a few thousand envs the GPU is almost empty (4096 envs = 128 warps on 592 SM sub-partitions)
// and the step time is K_max (~1340 substeps of the slowest env) x the time ONE warp needs per
// substep -- and that warp is bound by instruction issue: ~340 instructions per substep while the
// body shape moves (kinematics + dynamics + the fp64 shape chain and its ~90-instruction
// coefficient set), 166 afterwards.  But the substep is feed-forward:
//
//     front(j) -> coefs(j) -> dyn(j) -> kin(j)
//
//   * the shape and every coefficient derived from it depend on the action and on j only;
//   * the Newton/Euler equations + velocity update (dyn) need the coefficients and (v, w);
//   * the Euler angles / world position / body-frame integrals (kin) only consume (v, w) and never
//     feed back (there is no gravity or current in the reference's model).
// So one block of FOUR warps owns 32 envs, each warp on its own SM sub-partition:
//   * warp 2 (front) : shape_front(j), j = 1..W -- the fp64 shape chain and its backward
//                      differences, 8 floats per lane and substep into ring 1;
//   * warp 1 (coefs) : make_coefs(j) from ring 1 -- the stateless fp32 coefficient set, 28 floats
//                      per lane and substep into ring 2;
//   * warp 0 (dyn)   : the dyn recurrence, the only true critical path (~95 instructions per
//                      substep); coefficients from ring 2 while the shape moves, (v, w) into ring 3;
//   * warp 3 (kin)   : integrates the kinematics behind it from ring 3.
// Hand-off is chunk-granular (8 substeps, 2-3 chunks in flight per ring) on named barriers:
// bar.arrive on the side that is done with a chunk, bar.sync on the side that needs it, so no warp
// waits unless its neighbour has fallen a whole chunk behind.  The warps execute the functions of
// run_cycle_mixed (same fixed 32-substep grouping of the fp32 chunk sums); results agree with the
// fused kernel to fp32 rounding (the compiler contracts a*b+c differently in the two kernels;
// tests/test_gpu_parity.py::test_pipeline_kernel_matches_fused_kernel).
#pragma once
#include "salp_env.cuh"

#define SALP_PIPE_CHUNK 8
#define SALP_PIPE_NBUF1 2
#define SALP_PIPE_NBUF2 2
#define SALP_PIPE_NBUF3 3
#define SALP_PIPE_SLOTS1 (SALP_PIPE_CHUNK * SALP_PIPE_NBUF1)
#define SALP_PIPE_SLOTS2 (SALP_PIPE_CHUNK * SALP_PIPE_NBUF2)
#define SALP_PIPE_SLOTS3 32      // substep k in slot k % 32: at most 3 chunks (+ substep 0) = 25 entries in flight
#define SALP_PIPE_NCOEF 28
#define SALP_PIPE_THREADS 128

struct PipeShared {
  float ring2[SALP_PIPE_SLOTS2][32][SALP_PIPE_NCOEF];   // Coef32 of substep j in slot j % SLOTS2, one 112-byte row per lane
  float ring1[SALP_PIPE_SLOTS1][32][8];                 // ShapeFront of substep j in slot j % SLOTS1
  float ring3[SALP_PIPE_SLOTS3][32][8];                 // (v, w) after dyn(k) in slot k % SLOTS3
  double merge[22][32];                                 // final shape (front) and kinematic (kin) state, for warp 0's epilogue
};
static inline size_t pipe_smem_bytes(const SalpParams& p) {
  return sizeof(PipeShared) + sizeof(float) * 2 * 32 * (SALP_OBS_BASE + 2 * p.num_obstacles);
}

// named barriers (0 is __syncthreads); every hand-off involves two warps = 64 threads
__device__ __forceinline__ void pipe_bar_sync(int id) {
  asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
}
// (no fence: a completed barrier orders the shared-memory accesses its participants made before
//  arriving -- the producer/consumer idiom of the PTX ISA's bar.arrive / bar.sync example)
__device__ __forceinline__ void pipe_bar_arrive(int id) {
  asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory");
}
#define PIPE_FULL1(b) (1 + (b))
#define PIPE_EMPTY1(b) (1 + SALP_PIPE_NBUF1 + (b))
#define PIPE_FULL2(b) (1 + 2 * SALP_PIPE_NBUF1 + (b))
#define PIPE_EMPTY2(b) (1 + 2 * SALP_PIPE_NBUF1 + SALP_PIPE_NBUF2 + (b))
#define PIPE_FULL3(b) (1 + 2 * SALP_PIPE_NBUF1 + 2 * SALP_PIPE_NBUF2 + (b))
#define PIPE_EMPTY3(b) (1 + 2 * SALP_PIPE_NBUF1 + 2 * SALP_PIPE_NBUF2 + SALP_PIPE_NBUF3 + (b))

__device__ __forceinline__ void front_store(const ShapeFront& f, float* row) {
  float4* q = reinterpret_cast<float4*>(row);
  q[0] = make_float4(f.dl, f.I_rate0, f.I_rate1, f.dV_dt);
  q[1] = make_float4(f.com, f.com_rate, f.com_acc, f.jet_on);
}
__device__ __forceinline__ void front_load(ShapeFront& f, const float* row) {
  const float4* q = reinterpret_cast<const float4*>(row);
  float4 a = q[0], b = q[1];
  f.dl = a.x; f.I_rate0 = a.y; f.I_rate1 = a.z; f.dV_dt = a.w;
  f.com = b.x; f.com_rate = b.y; f.com_acc = b.z; f.jet_on = b.w;
}
__device__ __forceinline__ void vw_store(const Motion32& s, float* row) {
  float4* q = reinterpret_cast<float4*>(row);
  q[0] = make_float4(s.v0, s.v1, s.v2, s.w0);
  q[1] = make_float4(s.w1, s.w2, 0.f, 0.f);
}
__device__ __forceinline__ void vw_load(Motion32& s, const float* row) {
  const float4* q = reinterpret_cast<const float4*>(row);
  float4 a = q[0], b = q[1];
  s.v0 = a.x; s.v1 = a.y; s.v2 = a.z; s.w0 = a.w;
  s.w1 = b.x; s.w2 = b.y;
}
__device__ __forceinline__ void coef_store(const Coef32& g, float* row) {
  float4* q = reinterpret_cast<float4*>(row);
  q[0] = make_float4(g.aj[0], g.aj[1], g.aj[2], g.kdm[0]);
  q[1] = make_float4(g.kdm[1], g.kdm[2], g.mrm[0], g.mrm[1]);
  q[2] = make_float4(g.mrm[2], g.com, g.com_rate, g.com_acc);
  q[3] = make_float4(g.tj1, g.tj2, g.kqI[0], g.kqI[1]);
  q[4] = make_float4(g.kqI[2], g.klI[0], g.klI[1], g.klI[2]);
  q[5] = make_float4(g.JdI[0], g.JdI[1], g.JdI[2], g.AdI[0]);
  q[6] = make_float4(g.AdI[1], g.AdI[2], 0.f, 0.f);
}
__device__ __forceinline__ void coef_load(Coef32& g, const float* row) {
  const float4* q = reinterpret_cast<const float4*>(row);
  float4 a = q[0], b = q[1], c = q[2], d = q[3], e = q[4], f = q[5], h = q[6];
  g.aj[0] = a.x; g.aj[1] = a.y; g.aj[2] = a.z; g.kdm[0] = a.w;
  g.kdm[1] = b.x; g.kdm[2] = b.y; g.mrm[0] = b.z; g.mrm[1] = b.w;
  g.mrm[2] = c.x; g.com = c.y; g.com_rate = c.z; g.com_acc = c.w;
  g.tj1 = d.x; g.tj2 = d.y; g.kqI[0] = d.z; g.kqI[1] = d.w;
  g.kqI[2] = e.x; g.klI[0] = e.y; g.klI[1] = e.z; g.klI[2] = e.w;
  g.JdI[0] = f.x; g.JdI[1] = f.y; g.JdI[2] = f.z; g.AdI[0] = f.w;
  g.AdI[1] = h.x; g.AdI[2] = h.y;
}

__global__ void __launch_bounds__(SALP_PIPE_THREADS, 1)
salp_step_kernel_pipe(const __grid_constant__ SalpParams p, const __grid_constant__ SalpDerived dv,
                      const __grid_constant__ SalpView v, const __grid_constant__ SalpStepIO io, uint32_t flags) {
  extern __shared__ __align__(16) unsigned char pipe_smem[];
  PipeShared& sh = *reinterpret_cast<PipeShared*>(pipe_smem);
  float* tile = reinterpret_cast<float*>(pipe_smem + sizeof(PipeShared));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * 32 + lane;
  const bool live = i < v.n;
  constexpr int C = SALP_PIPE_CHUNK;

  // All three warps read the env's action and state themselves (reads only; every write happens in the
  // consumer's epilogue after the block-wide barrier) and derive the same integer plan.
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
  // body-frame integrals at the START of the cycle (env_step_end stores them as prev_position / prev_angle)
  const double pos0[3] = {b.pos[0], b.pos[1], b.pos[2]};
  const double ang0[3] = {b.ang[0], b.ang[1], b.ang[2]};
  // the same warp-uniform end of the shape-update part as run_cycle_mixed (updates j = 1..min(W, K))
  const int lane_end = pp.upd_a_end > pp.upd_b_end ? pp.upd_a_end : pp.upd_b_end;
  const int W = __reduce_max_sync(0xffffffffu, K > 0 ? (lane_end < K ? lane_end : K) : 0);
  const int Kw = __reduce_max_sync(0xffffffffu, K);
  const int kA = W < K ? W : K;
  const int Wmax = W < Kw ? W : Kw;
  const int nch = (Wmax + C - 1) / C;            // chunks of shape updates j = 1..Wmax
  const int nchK = Kw > 0 ? (Kw - 1 + C - 1) / C + (Kw == 1 ? 1 : 0) : 0;   // chunks of substeps 1..Kw-1 (one chunk if only substep 0 exists)
  const float dir[3] = {(float)cx.plan.dir[0], (float)cx.plan.dir[1], (float)cx.plan.dir[2]};

  if (warp == 2) {
    // ---------------- front: fp64 shape chain + backward differences, j = 1..kA ----------------
    ShapeTrack st;
    if (K > 0) {
      Coef32 g0;
      mixed_init_shape(p, dv, b, dir, st, g0);
    }
    double tj = v.time_table[1];                   // carried by the same additions as the table (robot.py:674)
    int j = 1;
    for (int c = 0; c < nch; c++) {
      if (c >= SALP_PIPE_NBUF1) pipe_bar_sync(PIPE_EMPTY1(c % SALP_PIPE_NBUF1));
      const int je = (c + 1) * C < Wmax ? (c + 1) * C : Wmax;
      // two updates per trip: consecutive updates are independent chains until their backward
      // differences (shape64_step carries nothing), so the scheduler overlaps them
      while (j <= je) {
        const double tj1 = rn::dadd(tj, p.dt);
        ShapeFront f0, f1;
        if (j + 1 <= je) {
          if (j + 1 <= kA) {
            shape_front(p, dv, cx.plan, tj, j, pp.k_T0, pp.k_jet, st, f0);
            shape_front(p, dv, cx.plan, tj1, j + 1, pp.k_T0, pp.k_jet, st, f1);
            front_store(f0, &sh.ring1[j % SALP_PIPE_SLOTS1][lane][0]);
            front_store(f1, &sh.ring1[(j + 1) % SALP_PIPE_SLOTS1][lane][0]);
          } else if (j <= kA) {
            shape_front(p, dv, cx.plan, tj, j, pp.k_T0, pp.k_jet, st, f0);
            front_store(f0, &sh.ring1[j % SALP_PIPE_SLOTS1][lane][0]);
          }
          tj = rn::dadd(tj1, p.dt);
          j += 2;
        } else {
          if (j <= kA) {
            shape_front(p, dv, cx.plan, tj, j, pp.k_T0, pp.k_jet, st, f0);
            front_store(f0, &sh.ring1[j % SALP_PIPE_SLOTS1][lane][0]);
          }
          tj = tj1;
          j += 1;
        }
      }
      __syncwarp();
      pipe_bar_arrive(PIPE_FULL1(c % SALP_PIPE_NBUF1));
    }
    if (K > 0) {
      mixed_finish_shape(p, st, K, b);
      sh.merge[0][lane] = b.length; sh.merge[1][lane] = b.width; sh.merge[2][lane] = b.prev_volume;
      sh.merge[3][lane] = b.prevI[0]; sh.merge[4][lane] = b.prevI[1];
      sh.merge[5][lane] = b.com; sh.merge[6][lane] = b.com_rate; sh.merge[7][lane] = b.prev_com_rate;
      sh.merge[8][lane] = b.com_acc;
    }
  } else if (warp == 1) {
    // ---------------- coefs: the stateless fp32 coefficient set of each ShapeFront ----------------
    int j = 1;
    for (int c = 0; c < nch; c++) {
      pipe_bar_sync(PIPE_FULL1(c % SALP_PIPE_NBUF1));
      if (c >= SALP_PIPE_NBUF2) pipe_bar_sync(PIPE_EMPTY2(c % SALP_PIPE_NBUF2));
      const int je = (c + 1) * C < Wmax ? (c + 1) * C : Wmax;
      while (j <= je) {
        ShapeFront f0, f1;
        Coef32 g0, g1;
        if (j + 1 <= je) {
          if (j + 1 <= kA) {
            front_load(f0, &sh.ring1[j % SALP_PIPE_SLOTS1][lane][0]);
            front_load(f1, &sh.ring1[(j + 1) % SALP_PIPE_SLOTS1][lane][0]);
            make_coefs(dv, dir, f0, g0);
            make_coefs(dv, dir, f1, g1);
            coef_store(g0, &sh.ring2[j % SALP_PIPE_SLOTS2][lane][0]);
            coef_store(g1, &sh.ring2[(j + 1) % SALP_PIPE_SLOTS2][lane][0]);
          } else if (j <= kA) {
            front_load(f0, &sh.ring1[j % SALP_PIPE_SLOTS1][lane][0]);
            make_coefs(dv, dir, f0, g0);
            coef_store(g0, &sh.ring2[j % SALP_PIPE_SLOTS2][lane][0]);
          }
          j += 2;
        } else {
          if (j <= kA) {
            front_load(f0, &sh.ring1[j % SALP_PIPE_SLOTS1][lane][0]);
            make_coefs(dv, dir, f0, g0);
            coef_store(g0, &sh.ring2[j % SALP_PIPE_SLOTS2][lane][0]);
          }
          j += 1;
        }
      }
      __syncwarp();
      pipe_bar_arrive(PIPE_EMPTY1(c % SALP_PIPE_NBUF1));
      pipe_bar_arrive(PIPE_FULL2(c % SALP_PIPE_NBUF2));
    }
  } else if (warp == 0) {
    // ---------------- dyn: the recurrence; substeps k = 0..K-1, chunk c = substeps 8c+1..8c+8 (+ substep 0 in chunk 0) ----------------
    Motion32 s;
    Coef32 g;
    if (K > 0) {
      ShapeTrack st0;
      mixed_init_shape(p, dv, b, dir, st0, g);      // g_0 (once; cheaper than a hand-off)
      mixed_init_dyn(b, s);
      dyn_step(dv, g, s);
      vw_store(s, &sh.ring3[0][lane][0]);
    }
    int kk = 1;
    for (int c = 0; c < nchK; c++) {
      if (c < nch) pipe_bar_sync(PIPE_FULL2(c % SALP_PIPE_NBUF2));
      if (c >= SALP_PIPE_NBUF3) pipe_bar_sync(PIPE_EMPTY3(c % SALP_PIPE_NBUF3));
      const int ce = (c + 1) * C < Kw - 1 ? (c + 1) * C : Kw - 1;
      if (c < nch) {
        for (; kk <= ce; kk++) {
          if (kk < K) {
            if (kk <= W) coef_load(g, &sh.ring2[kk % SALP_PIPE_SLOTS2][lane][0]);
            dyn_step(dv, g, s);
            vw_store(s, &sh.ring3[kk % SALP_PIPE_SLOTS3][lane][0]);
          }
        }
      } else {
        for (; kk <= ce; kk++) {
          if (kk < K) {
            dyn_step(dv, g, s);
            vw_store(s, &sh.ring3[kk % SALP_PIPE_SLOTS3][lane][0]);
          }
        }
      }
      __syncwarp();
      if (c < nch) pipe_bar_arrive(PIPE_EMPTY2(c % SALP_PIPE_NBUF2));
      pipe_bar_arrive(PIPE_FULL3(c % SALP_PIPE_NBUF3));
    }
    if (K > 0) mixed_finish_dyn(s, b);
  } else {
    // ---------------- kin: Euler angles, world position, body-frame integrals; kin(k) from (v, w) after dyn(k) ----------------
    Motion32 s;
    if (K > 0) mixed_init_kin(b, s);
    int kk = 0;
    for (int c = 0; c < nchK; c++) {
      pipe_bar_sync(PIPE_FULL3(c % SALP_PIPE_NBUF3));
      const int ce = (c + 1) * C < Kw - 1 ? (c + 1) * C : Kw - 1;
      // two substeps per trip where possible: the rotation of v into the world frame and the
      // integrals of substep k overlap the Euler-rate chain of substep k + 1
      // (kk and the trip structure stay warp-uniform; only the work inside is per lane)
      while (kk <= ce) {
        if (kk + 1 <= ce && ((kk + 1) & (SALP_MIXED_CHUNK - 1)) != 0) {       // no flush between the two
          if (kk + 1 < K) {
            vw_load(s, &sh.ring3[kk % SALP_PIPE_SLOTS3][lane][0]);
            kin_step(dv, s);
            vw_load(s, &sh.ring3[(kk + 1) % SALP_PIPE_SLOTS3][lane][0]);
            kin_step(dv, s);
            // the fused loop flushes after iteration 32 m (kinematic updates 0..32 m - 1 done) if 32 m < K
            if (((kk + 2) & (SALP_MIXED_CHUNK - 1)) == 0 && kk + 2 < K) flush_chunk(b, s);
          } else if (kk < K) {
            vw_load(s, &sh.ring3[kk % SALP_PIPE_SLOTS3][lane][0]);
            kin_step(dv, s);
          }
          kk += 2;
        } else {
          if (kk < K) {
            vw_load(s, &sh.ring3[kk % SALP_PIPE_SLOTS3][lane][0]);
            kin_step(dv, s);
            if (((kk + 1) & (SALP_MIXED_CHUNK - 1)) == 0 && kk + 1 < K) flush_chunk(b, s);
          }
          kk += 1;
        }
      }
      __syncwarp();
      pipe_bar_arrive(PIPE_EMPTY3(c % SALP_PIPE_NBUF3));
    }
    if (K > 0) {
      flush_chunk(b, s);
#pragma unroll
      for (int k = 0; k < 3; k++) {
        sh.merge[9 + k][lane] = b.pw[k]; sh.merge[12 + k][lane] = b.pos[k];
        sh.merge[15 + k][lane] = b.ang[k]; sh.merge[18 + k][lane] = b.eul[k];
      }
      sh.merge[21][lane] = (double)sqrtf(s.vw0 * s.vw0 + s.vw1 * s.vw1);
    }
  }
  __syncthreads();
  if (warp != 0) return;
  if (live) {
    double t = 0.0;
    if (K > 0) {
      b.length = sh.merge[0][lane]; b.width = sh.merge[1][lane]; b.prev_volume = sh.merge[2][lane];
      b.prevI[0] = sh.merge[3][lane]; b.prevI[1] = sh.merge[4][lane]; b.prevI[2] = sh.merge[4][lane];
      b.com = sh.merge[5][lane]; b.prev_com = sh.merge[5][lane]; b.com_rate = sh.merge[6][lane];
      b.prev_com_rate = sh.merge[7][lane]; b.com_acc = sh.merge[8][lane];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        b.pw[k] = sh.merge[9 + k][lane]; b.pos[k] = sh.merge[12 + k][lane];
        b.ang[k] = sh.merge[15 + k][lane]; b.eul[k] = sh.merge[18 + k][lane];
      }
      b.speed_world = sh.merge[21][lane];
      t = v.time_table[K];
      b.phase = phase_at(cx.plan, t);
    }
    env_step_end(p, v, io, flags, i, cx, pos0, ang0, b, Kraw, t, tile + lane * (SALP_OBS_BASE + 2 * p.num_obstacles),
                 io.terminal_obs ? tile + (32 + lane) * (SALP_OBS_BASE + 2 * p.num_obstacles) : nullptr);
  }
  __syncwarp();
  const int D = SALP_OBS_BASE + 2 * p.num_obstacles;
  const int rows = __popc(__ballot_sync(0xffffffffu, live));
  for (int j = lane; j < 32 * D; j += 32) {
    if (j < rows * D) {
      io.obs[(int64_t)blockIdx.x * 32 * D + j] = tile[j];
      if (io.terminal_obs) io.terminal_obs[(int64_t)blockIdx.x * 32 * D + j] = tile[32 * D + j];
    }
  }
}
