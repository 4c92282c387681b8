// salp_pipe4_kernel.cuh -- the small-batch step kernel, four warps per 32 envs.
//
// With a few thousand envs the GPU is almost empty and one env-step costs K_max (~1340 substeps of
// the slowest env) x the time ONE warp needs per substep; that warp is bound by instruction issue
// (ncu: the three-warp kernel's consumer issues on ~80 % of its scheduler's cycles).  So the
// substep is cut along its feed-forward structure into four instruction streams, one warp each,
// each warp on its own SM sub-partition:
//
//   warp 3 (front) : fp64 shape chain + backward differences of substep j      -> ring 1  (8 floats / lane)
//   warp 2 (coefs) : the stateless fp32 coefficient set from ring 1            -> ring 2  (20 or 28 floats)
//   warp 0 (dyn)   : the ONLY recurrence that feeds back: (v, w, a, alpha) of substep j from ring 2,
//                    plus the body-frame integrals (kin_body: they only need v, w)  -> ring 3  (v, w: 6 floats)
//   warp 1 (kin)   : Euler angles and world position from the (v, w) stream (kin_world); nothing
//                    flows back to the dynamics, so this warp simply trails the dyn warp.
//
// The shape and every coefficient depend on the action and the substep index only; the kinematics
// depend on (v, w) only.  Hand-off is chunk-granular (8 substeps) on named barriers -- bar.arrive by
// the side that is done with a chunk, bar.sync by the side that needs it -- with 2 / 3 / 2 chunks
// in flight on rings 1 / 2 / 3 (14 of the 16 hardware barriers).  The warps execute the functions of
// run_cycle_mixed with the same fixed 32-substep grouping of the fp32 chunk sums, and every
// operation of the loop is explicitly rounded: bit-identical with the fused kernel and the
// three-warp kernel (tests/test_gpu_parity.py).
//
// Envs may be visited through a permutation (`order`, the K-sort), and two blocks fit an SM
// (<= 110 KB of rings each), so the kernel also serves batches of up to 64 envs per SM.
#pragma once
#include "salp_pipe_kernel.cuh"

#define SALP_P4_CHUNK 8
#define SALP_P4_NBUF1 2
#define SALP_P4_NBUF2 3
#define SALP_P4_NBUF3 2
#define SALP_P4_SLOTS1 (SALP_P4_CHUNK * SALP_P4_NBUF1)
#define SALP_P4_SLOTS2 (SALP_P4_CHUNK * SALP_P4_NBUF2)
#define SALP_P4_SLOTS3 (SALP_P4_CHUNK * SALP_P4_NBUF3)
#define SALP_P4_THREADS 128
#define SALP_P4_MERGE 16

#define P4_FULL1(b) (1 + (b))
#define P4_EMPTY1(b) (1 + SALP_P4_NBUF1 + (b))
#define P4_FULL2(b) (1 + 2 * SALP_P4_NBUF1 + (b))
#define P4_EMPTY2(b) (1 + 2 * SALP_P4_NBUF1 + SALP_P4_NBUF2 + (b))
#define P4_FULL3(b) (1 + 2 * SALP_P4_NBUF1 + 2 * SALP_P4_NBUF2 + (b))
#define P4_EMPTY3(b) (1 + 2 * SALP_P4_NBUF1 + 2 * SALP_P4_NBUF2 + SALP_P4_NBUF3 + (b))

// dynamic shared memory: ring 2 rows are 80 bytes (axisymmetric form) or 112 bytes per lane
template <bool AXI>
struct P4Layout {
  static constexpr int ROW = AXI ? 20 : SALP_PIPE_NCOEF;
  static constexpr size_t ring2 = 0;                                                          // [SLOTS2][32][ROW] float
  static constexpr size_t ring1 = ring2 + sizeof(float) * SALP_P4_SLOTS2 * 32 * ROW;          // [SLOTS1][32][8] float
  static constexpr size_t ring3a = ring1 + sizeof(float) * SALP_P4_SLOTS1 * 32 * 8;           // [SLOTS3][32] float4 (v0 v1 v2 w0)
  static constexpr size_t ring3b = ring3a + sizeof(float4) * SALP_P4_SLOTS3 * 32;             // [SLOTS3][32] float2 (w1 w2)
  static constexpr size_t merge = ring3b + sizeof(float2) * SALP_P4_SLOTS3 * 32;              // [MERGE][32] double
  static constexpr size_t tile = merge + sizeof(double) * SALP_P4_MERGE * 32;                 // [2][32][D] float
};
static inline size_t pipe4_smem_bytes(const SalpParams& p, bool axi) {
  const size_t tile = sizeof(float) * 2 * 32 * (SALP_OBS_BASE + 2 * p.num_obstacles);
  return (axi ? P4Layout<true>::tile : P4Layout<false>::tile) + tile;
}

template <bool AXI>
__device__ __forceinline__ void salp_pipe4_body(const SalpParams& p, const SalpDerived& dv, const SalpView& v,
                                                const SalpStepIO& io, uint32_t flags, const int32_t* __restrict__ order,
                                                unsigned char* smem) {
  using L = P4Layout<AXI>;
  constexpr int ROW = L::ROW;
  constexpr int C = SALP_P4_CHUNK;
  float* ring2 = reinterpret_cast<float*>(smem + L::ring2);
  float* ring1 = reinterpret_cast<float*>(smem + L::ring1);
  float4* ring3a = reinterpret_cast<float4*>(smem + L::ring3a);
  float2* ring3b = reinterpret_cast<float2*>(smem + L::ring3b);
  double* merge = reinterpret_cast<double*>(smem + L::merge);
  float* tile = reinterpret_cast<float*>(smem + L::tile);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t tid = (int64_t)blockIdx.x * 32 + lane;
  const bool live = tid < v.n;
  const int64_t i = live ? (order ? (int64_t)order[tid] : tid) : 0;

  // All four warps read the env's action and state themselves (reads only; every write happens in
  // the dyn warp's epilogue after the block-wide barrier) and derive the same integer plan.
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
  const double pos0[3] = {b.pos[0], b.pos[1], b.pos[2]};
  const double ang0[3] = {b.ang[0], b.ang[1], b.ang[2]};
  // the same warp-uniform end of the shape-update part as run_cycle_mixed (updates j = 1..min(W, K))
  const int lane_end = pp.upd_a_end > pp.upd_b_end ? pp.upd_a_end : pp.upd_b_end;
  const int W = __reduce_max_sync(0xffffffffu, K > 0 ? (lane_end < K ? lane_end : K) : 0);
  const int Kw = __reduce_max_sync(0xffffffffu, K);
  const int kA = W < K ? W : K;                      // this lane's last shape update
  const int Wmax = W < Kw ? W : Kw;                  // the producers run updates j = 1..Wmax
  const int WA = Wmax < Kw - 1 ? Wmax : Kw - 1;      // substeps j = 1..K-1 exist; those <= W use g_j
  const int nch2 = (Wmax + C - 1) / C;               // chunk c = substeps c C + 1 .. (c + 1) C
  const int nch3 = Kw > 1 ? (Kw - 1 + C - 1) / C : 0;
  const float dir[3] = {(float)cx.plan.dir[0], (float)cx.plan.dir[1], (float)cx.plan.dir[2]};

  if (warp == 3) {
    // ---------------- front: fp64 shape chain + backward differences, j = 1..kA ----------------
    ShapeTrack st;
    if (K > 0) {
      Coef32 g0;
      mixed_init_shape<AXI>(p, dv, b, dir, st, g0);
    }
    double tj = v.time_table[1];                   // carried by the same additions as the table (robot.py:674)
    int j = 1;
    for (int c = 0; c < nch2; c++) {
      if (c >= SALP_P4_NBUF1) pipe_bar_sync(P4_EMPTY1(c % SALP_P4_NBUF1));
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
            front_store(f0, ring1 + ((j % SALP_P4_SLOTS1) * 32 + lane) * 8);
            front_store(f1, ring1 + (((j + 1) % SALP_P4_SLOTS1) * 32 + lane) * 8);
          } else if (j <= kA) {
            shape_front(p, dv, cx.plan, tj, j, pp.k_T0, pp.k_jet, st, f0);
            front_store(f0, ring1 + ((j % SALP_P4_SLOTS1) * 32 + lane) * 8);
          }
          tj = rn::dadd(tj1, p.dt);
          j += 2;
        } else {
          if (j <= kA) {
            shape_front(p, dv, cx.plan, tj, j, pp.k_T0, pp.k_jet, st, f0);
            front_store(f0, ring1 + ((j % SALP_P4_SLOTS1) * 32 + lane) * 8);
          }
          tj = tj1;
          j += 1;
        }
      }
      __syncwarp();
      pipe_bar_arrive(P4_FULL1(c % SALP_P4_NBUF1));
    }
    if (K > 0) {
      mixed_finish_shape(p, st, K, b);
      merge[0 * 32 + lane] = b.length; merge[1 * 32 + lane] = b.width; merge[2 * 32 + lane] = b.prev_volume;
      merge[3 * 32 + lane] = b.prevI[0]; merge[4 * 32 + lane] = b.prevI[1];
      merge[5 * 32 + lane] = b.com; merge[6 * 32 + lane] = b.com_rate; merge[7 * 32 + lane] = b.prev_com_rate;
      merge[8 * 32 + lane] = b.com_acc;
    }
  } else if (warp == 2) {
    // ---------------- coefs: the stateless fp32 coefficient set of each ShapeFront ----------------
    int j = 1;
    for (int c = 0; c < nch2; c++) {
      pipe_bar_sync(P4_FULL1(c % SALP_P4_NBUF1));
      if (c >= SALP_P4_NBUF2) pipe_bar_sync(P4_EMPTY2(c % SALP_P4_NBUF2));
      const int je = (c + 1) * C < Wmax ? (c + 1) * C : Wmax;
      while (j <= je) {
        ShapeFront f0, f1;
        Coef32 g0, g1;
        if (j + 1 <= je) {
          if (j + 1 <= kA) {
            front_load(f0, ring1 + ((j % SALP_P4_SLOTS1) * 32 + lane) * 8);
            front_load(f1, ring1 + (((j + 1) % SALP_P4_SLOTS1) * 32 + lane) * 8);
            make_coefs<AXI>(dv, dir, f0, g0);
            make_coefs<AXI>(dv, dir, f1, g1);
            coef_store<AXI>(g0, ring2 + ((j % SALP_P4_SLOTS2) * 32 + lane) * ROW);
            coef_store<AXI>(g1, ring2 + (((j + 1) % SALP_P4_SLOTS2) * 32 + lane) * ROW);
          } else if (j <= kA) {
            front_load(f0, ring1 + ((j % SALP_P4_SLOTS1) * 32 + lane) * 8);
            make_coefs<AXI>(dv, dir, f0, g0);
            coef_store<AXI>(g0, ring2 + ((j % SALP_P4_SLOTS2) * 32 + lane) * ROW);
          }
          j += 2;
        } else {
          if (j <= kA) {
            front_load(f0, ring1 + ((j % SALP_P4_SLOTS1) * 32 + lane) * 8);
            make_coefs<AXI>(dv, dir, f0, g0);
            coef_store<AXI>(g0, ring2 + ((j % SALP_P4_SLOTS2) * 32 + lane) * ROW);
          }
          j += 1;
        }
      }
      __syncwarp();
      pipe_bar_arrive(P4_EMPTY1(c % SALP_P4_NBUF1));
      pipe_bar_arrive(P4_FULL2(c % SALP_P4_NBUF2));
    }
  } else if (warp == 1) {
    // ---------------- kin: Euler angles and world position from the (v, w) stream ----------------
    // kin step j uses (v, w) after dyn(j).  Step 0's input is recomputed here (g_0 and dyn(0): cheaper
    // than a one-off hand-off), steps 1..K-1 come through ring 3.  The lane's LAST step and the final
    // flush run converged after the loop, like in the fused kernel.
    Motion32 s;
    if (K > 0) {
      ShapeTrack st0;
      Coef32 g;
      mixed_init_shape<AXI>(p, dv, b, dir, st0, g);
      mixed_init_dyn(b, s);
      mixed_init_kin(b, s);
      dyn_step<false, false, AXI>(dv, g, s);
    }
    for (int c = 0; c < nch3; c++) {
      pipe_bar_sync(P4_FULL3(c % SALP_P4_NBUF3));
      const int je = (c + 1) * C < Kw - 1 ? (c + 1) * C : Kw - 1;
#pragma unroll 1
      for (int j = c * C + 1; j <= je; j++) {
        if (j < K) {
          kin_world(dv, s);                                     // step j - 1, on the (v, w) loaded one trip earlier
          if ((j & (SALP_MIXED_CHUNK - 1)) == 0) flush_world(b, s);
          const float4 a = ring3a[(j % SALP_P4_SLOTS3) * 32 + lane];
          const float2 bb = ring3b[(j % SALP_P4_SLOTS3) * 32 + lane];
          s.v0 = a.x; s.v1 = a.y; s.v2 = a.z; s.w0 = a.w; s.w1 = bb.x; s.w2 = bb.y;
        }
      }
      __syncwarp();
      pipe_bar_arrive(P4_EMPTY3(c % SALP_P4_NBUF3));
    }
    if (K > 0) {
      kin_world(dv, s);                                         // step K - 1
      flush_world(b, s);
      merge[9 * 32 + lane] = b.eul[0]; merge[10 * 32 + lane] = b.eul[1]; merge[11 * 32 + lane] = b.eul[2];
      merge[12 * 32 + lane] = b.pw[0]; merge[13 * 32 + lane] = b.pw[1]; merge[14 * 32 + lane] = b.pw[2];
      merge[15 * 32 + lane] = (double)sqrtf(s.vw0 * s.vw0 + s.vw1 * s.vw1);
    }
  } else {
    // ---------------- dyn: the (v, w, a, alpha) recurrence + body-frame integrals ----------------
    Motion32 s;
    Coef32 g;
    if (K > 0) {
      ShapeTrack st0;
      mixed_init_shape<AXI>(p, dv, b, dir, st0, g);      // g_0 (once; cheaper than a hand-off)
      mixed_init_dyn(b, s);
      s.pos0 = s.pos1 = s.pos2 = 0.f; s.ang0 = s.ang1 = s.ang2 = 0.f;
      dyn_step<false, false, AXI>(dv, g, s);
    }
    const int nch = nch2 > nch3 ? nch2 : nch3;
    for (int c = 0; c < nch; c++) {
      if (c < nch2) pipe_bar_sync(P4_FULL2(c % SALP_P4_NBUF2));
      if (c < nch3 && c >= SALP_P4_NBUF3) pipe_bar_sync(P4_EMPTY3(c % SALP_P4_NBUF3));
      const int je = (c + 1) * C < Kw - 1 ? (c + 1) * C : Kw - 1;
      if (c * C + 1 <= WA) {
        // the shape is (or may still be) moving somewhere in this chunk: coefficients from ring 2
#pragma unroll 1
        for (int j = c * C + 1; j <= je; j++) {
          if (j < K) {
            if (j <= WA) coef_load<AXI>(g, ring2 + ((j % SALP_P4_SLOTS2) * 32 + lane) * ROW);
            kin_body(dv, s);                                    // step j - 1
            dyn_step<false, false, AXI>(dv, g, s);              // (j > W: com_rate = com_acc = 0, same bits as the static form)
            ring3a[(j % SALP_P4_SLOTS3) * 32 + lane] = make_float4(s.v0, s.v1, s.v2, s.w0);
            ring3b[(j % SALP_P4_SLOTS3) * 32 + lane] = make_float2(s.w1, s.w2);
            if ((j & (SALP_MIXED_CHUNK - 1)) == 0) flush_body(b, s);
          }
        }
      } else {
        // the coast: static shape, coefficients stay in registers
#pragma unroll 1
        for (int j = c * C + 1; j <= je; j++) {
          if (j < K) {
            kin_body(dv, s);
            dyn_step<false, true, AXI>(dv, g, s);
            ring3a[(j % SALP_P4_SLOTS3) * 32 + lane] = make_float4(s.v0, s.v1, s.v2, s.w0);
            ring3b[(j % SALP_P4_SLOTS3) * 32 + lane] = make_float2(s.w1, s.w2);
            if ((j & (SALP_MIXED_CHUNK - 1)) == 0) flush_body(b, s);
          }
        }
      }
      __syncwarp();
      if (c < nch2) pipe_bar_arrive(P4_EMPTY2(c % SALP_P4_NBUF2));
      if (c < nch3) pipe_bar_arrive(P4_FULL3(c % SALP_P4_NBUF3));
    }
    if (K > 0) {
      kin_body(dv, s);                                          // step K - 1
      flush_body(b, s);
      mixed_finish_dyn(s, b);
    }
  }
  __syncthreads();
  if (warp != 0) return;
  const int D = SALP_OBS_BASE + 2 * p.num_obstacles;
  if (live) {
    double t = 0.0;
    if (K > 0) {
      b.length = merge[0 * 32 + lane]; b.width = merge[1 * 32 + lane]; b.prev_volume = merge[2 * 32 + lane];
      b.prevI[0] = merge[3 * 32 + lane]; b.prevI[1] = merge[4 * 32 + lane]; b.prevI[2] = merge[4 * 32 + lane];
      b.com = merge[5 * 32 + lane]; b.prev_com = merge[5 * 32 + lane]; b.com_rate = merge[6 * 32 + lane];
      b.prev_com_rate = merge[7 * 32 + lane]; b.com_acc = merge[8 * 32 + lane];
      b.eul[0] = merge[9 * 32 + lane]; b.eul[1] = merge[10 * 32 + lane]; b.eul[2] = merge[11 * 32 + lane];
      b.pw[0] = merge[12 * 32 + lane]; b.pw[1] = merge[13 * 32 + lane]; b.pw[2] = merge[14 * 32 + lane];
      b.speed_world = merge[15 * 32 + lane];
      t = v.time_table[K];
      b.phase = phase_at(cx.plan, t);
    }
    env_step_end(p, v, io, flags, i, cx, pos0, ang0, b, Kraw, t, tile + lane * D,
                 io.terminal_obs ? tile + (32 + lane) * D : nullptr);
  }
  __syncwarp();
  const int rows = __popc(__ballot_sync(0xffffffffu, live));     // live lanes are the low lanes
  for (int j = lane; j < 32 * D; j += 32) {                      // warp-uniform trip count (D iterations)
    const int r = j / D, k = j - r * D;
    const int64_t e = __shfl_sync(0xffffffffu, i, r);
    if (r < rows) {
      io.obs[e * D + k] = tile[j];
      if (io.terminal_obs) io.terminal_obs[e * D + k] = tile[32 * D + j];
    }
  }
}

__global__ void __launch_bounds__(SALP_P4_THREADS, 2)
salp_step_kernel_pipe4(const __grid_constant__ SalpParams p, const __grid_constant__ SalpDerived dv,
                       const __grid_constant__ SalpView v, const __grid_constant__ SalpStepIO io, uint32_t flags,
                       const int32_t* __restrict__ order) {
  extern __shared__ __align__(16) unsigned char pipe4_smem[];
  // (block-uniform: dv is a kernel argument; both forms give the same bits for axisymmetric parameters)
  if (dv.axisym) salp_pipe4_body<true>(p, dv, v, io, flags, order, pipe4_smem);
  else salp_pipe4_body<false>(p, dv, v, io, flags, order, pipe4_smem);
}
