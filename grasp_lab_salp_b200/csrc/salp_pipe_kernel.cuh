// salp_pipe_kernel.cuh -- the small-batch step kernel: two shape-producer warps + one motion warp.
//
// With a few thousand envs the GPU is almost empty (4096 envs = 128 warps on 592 SM sub-partitions)
// and the step time is K_max (~1340 substeps of the slowest env) x the time ONE warp needs per
// substep -- and that warp is bound by instruction issue: ~330 instructions per substep while the
// body shape moves (kinematics + dynamics + the fp64 shape chain and its ~90-instruction
// coefficient set), ~150 afterwards.  But the shape and every coefficient derived from it depend on
// the action and the substep index only, never on the motion state.  So one block of THREE warps
// owns 32 envs, each warp on its own SM sub-partition:
//   * warp 2 (front)    : shape_front(j), j = 1..W -- the fp64 shape chain and its backward
//                         differences, 8 floats per lane and substep into ring 1;
//   * warp 1 (coefs)    : make_coefs(j) from ring 1 -- the stateless fp32 coefficient set, 28 floats
//                         per lane and substep into ring 2;
//   * warp 0 (consumer) : the same software-pipelined kin(k-1) || dyn(k) loop as the fused kernel,
//                         loading its coefficients from ring 2 instead of computing them.
// Hand-off is chunk-granular (8 substeps; 3 resp. 4 chunks in flight) on named barriers:
// bar.arrive on the side that is done with a chunk, bar.sync on the side that needs it, so no warp
// waits unless its neighbour has fallen a whole chunk behind.  The three warps execute the
// functions of run_cycle_mixed (same fixed 32-substep grouping of the fp32 chunk sums, every
// operation of the loop explicitly rounded): results are bit-identical with the fused kernel
// (tests/test_gpu_parity.py::test_pipeline_kernel_matches_fused_kernel).
// (Splitting the consumer further into a dyn warp and a kin warp was measured slower, 0.203 vs
// 0.190 ms per 4096-env step: in one warp the two chains fill each other's latency shadows.)
#pragma once
#include "salp_env.cuh"

#define SALP_PIPE_CHUNK 8
#define SALP_PIPE_NBUF1 3
#define SALP_PIPE_NBUF2 4
#define SALP_PIPE_SLOTS1 (SALP_PIPE_CHUNK * SALP_PIPE_NBUF1)
#define SALP_PIPE_SLOTS2 (SALP_PIPE_CHUNK * SALP_PIPE_NBUF2)
#define SALP_PIPE_NCOEF 28
#define SALP_PIPE_THREADS 96

struct PipeShared {
  float ring2[SALP_PIPE_SLOTS2][32][SALP_PIPE_NCOEF];   // Coef32 of substep j in slot j % SLOTS2, one 112-byte row per lane
  float ring1[SALP_PIPE_SLOTS1][32][8];                 // ShapeFront of substep j in slot j % SLOTS1
  double merge[9][32];                                  // the front warp's final shape state, for the consumer's epilogue
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
// One 112-byte row per lane and slot.  The first five quads are everything the axisymmetric form
// needs (AXI: two fewer 128-bit accesses per substep on each side); the last two hold the entries
// that only differ for asymmetric coefficient sets.
template <bool AXI>
__device__ __forceinline__ void coef_store(const Coef32& g, float* row) {
  float4* q = reinterpret_cast<float4*>(row);
  q[0] = make_float4(g.aj[0], g.aj[1], g.aj[2], g.kdm[0]);
  q[1] = make_float4(g.kdm[1], g.xc[0], g.com, g.com_rate2);
  q[2] = make_float4(g.com_acc, g.tj1, g.tj2, g.kqI[0]);
  q[3] = make_float4(g.kqI[1], g.klI[0], g.klI[1], g.JdI[1]);
  if (AXI) {
    q[4] = make_float4(g.AdI[1], g.xc[1], 0.f, 0.f);
  } else {
    q[4] = make_float4(g.AdI[1], g.xc[1], g.kdm[2], g.xc[2]);
    q[5] = make_float4(g.kqI[2], g.klI[2], g.JdI[0], g.JdI[2]);
    q[6] = make_float4(g.AdI[0], g.AdI[2], 0.f, 0.f);
  }
}
template <bool AXI>
__device__ __forceinline__ void coef_load(Coef32& g, const float* row) {
  const float4* q = reinterpret_cast<const float4*>(row);
  float4 a = q[0], b = q[1], c = q[2], d = q[3], e = q[4];
  g.aj[0] = a.x; g.aj[1] = a.y; g.aj[2] = a.z; g.kdm[0] = a.w;
  g.kdm[1] = b.x; g.xc[0] = b.y; g.com = b.z; g.com_rate2 = b.w;
  g.com_acc = c.x; g.tj1 = c.y; g.tj2 = c.z; g.kqI[0] = c.w;
  g.kqI[1] = d.x; g.klI[0] = d.y; g.klI[1] = d.z; g.JdI[1] = d.w;
  g.AdI[1] = e.x; g.xc[1] = e.y;
  if (!AXI) {
    float4 f = q[5], h = q[6];
    g.kdm[2] = e.z; g.xc[2] = e.w;
    g.kqI[2] = f.x; g.klI[2] = f.y; g.JdI[0] = f.z; g.JdI[2] = f.w;
    g.AdI[0] = h.x; g.AdI[2] = h.y;
  }
}

template <bool AXI>
__device__ __forceinline__ void salp_pipe_body(const SalpParams& p, const SalpDerived& dv, const SalpView& v,
                                               const SalpStepIO& io, uint32_t flags, unsigned char* pipe_smem) {
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
  const int nch = (Wmax + C - 1) / C;
  const float dir[3] = {(float)cx.plan.dir[0], (float)cx.plan.dir[1], (float)cx.plan.dir[2]};

  if (warp == 2) {
    // ---------------- front: fp64 shape chain + backward differences, j = 1..kA ----------------
    ShapeTrack st;
    if (K > 0) {
      Coef32 g0;
      mixed_init_shape<AXI>(p, dv, b, dir, st, g0);
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
            make_coefs<AXI>(dv, dir, f0, g0);
            make_coefs<AXI>(dv, dir, f1, g1);
            coef_store<AXI>(g0, &sh.ring2[j % SALP_PIPE_SLOTS2][lane][0]);
            coef_store<AXI>(g1, &sh.ring2[(j + 1) % SALP_PIPE_SLOTS2][lane][0]);
          } else if (j <= kA) {
            front_load(f0, &sh.ring1[j % SALP_PIPE_SLOTS1][lane][0]);
            make_coefs<AXI>(dv, dir, f0, g0);
            coef_store<AXI>(g0, &sh.ring2[j % SALP_PIPE_SLOTS2][lane][0]);
          }
          j += 2;
        } else {
          if (j <= kA) {
            front_load(f0, &sh.ring1[j % SALP_PIPE_SLOTS1][lane][0]);
            make_coefs<AXI>(dv, dir, f0, g0);
            coef_store<AXI>(g0, &sh.ring2[j % SALP_PIPE_SLOTS2][lane][0]);
          }
          j += 1;
        }
      }
      __syncwarp();
      pipe_bar_arrive(PIPE_EMPTY1(c % SALP_PIPE_NBUF1));
      pipe_bar_arrive(PIPE_FULL2(c % SALP_PIPE_NBUF2));
    }
  } else {
    // ---------------- consumer: kin(k-1) || dyn(k), coefficients from the ring ----------------
    Motion32 s;
    Coef32 g;
    if (K > 0) {
      ShapeTrack st0;
      mixed_init_shape<AXI>(p, dv, b, dir, st0, g);      // g_0 (once; cheaper than a hand-off)
      mixed_init_dyn(b, s);
      mixed_init_kin(b, s);
      dyn_step<false, false, AXI>(dv, g, s);
    }
    int kk = 1;
    const int WA = Wmax < Kw - 1 ? Wmax : Kw - 1;   // iterations kk = 1..K-1 exist; those <= W load g_kk
    for (int c = 0; c < nch; c++) {
      pipe_bar_sync(PIPE_FULL2(c % SALP_PIPE_NBUF2));
      const int ce = (c + 1) * C < WA ? (c + 1) * C : WA;
      for (; kk <= ce; kk++) {
        if (kk < K) {
          coef_load<AXI>(g, &sh.ring2[kk % SALP_PIPE_SLOTS2][lane][0]);
          kin_step(dv, s);
          dyn_step<false, false, AXI>(dv, g, s);
          if ((kk & (SALP_MIXED_CHUNK - 1)) == 0) flush_chunk(b, s);
        }
      }
      __syncwarp();
      pipe_bar_arrive(PIPE_EMPTY2(c % SALP_PIPE_NBUF2));
    }
    // the coast: the fused kernel's lean loop, same fixed chunk boundaries
    int k = kk;
    while (k < K) {
      const int boundary = ((k - 1) & ~(SALP_MIXED_CHUNK - 1)) + SALP_MIXED_CHUNK + 1;
      const int cend = boundary < K ? boundary : K;
      for (; k < cend; k++) {
        kin_step(dv, s);
        dyn_step<false, true, AXI>(dv, g, s);       // k > W: the shape is static
      }
      if (k == boundary) flush_chunk(b, s);
    }
    if (K > 0) {
      kin_step(dv, s);
      flush_chunk(b, s);
      mixed_finish_dyn(s, b);
      b.speed_world = (double)sqrtf(s.vw0 * s.vw0 + s.vw1 * s.vw1);
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

__global__ void __launch_bounds__(SALP_PIPE_THREADS, 1)
salp_step_kernel_pipe(const __grid_constant__ SalpParams p, const __grid_constant__ SalpDerived dv,
                      const __grid_constant__ SalpView v, const __grid_constant__ SalpStepIO io, uint32_t flags) {
  extern __shared__ __align__(16) unsigned char pipe_smem[];
  // (block-uniform: dv is a kernel argument; both forms give the same bits for axisymmetric parameters)
  if (dv.axisym) salp_pipe_body<true>(p, dv, v, io, flags, pipe_smem);
  else salp_pipe_body<false>(p, dv, v, io, flags, pipe_smem);
}
