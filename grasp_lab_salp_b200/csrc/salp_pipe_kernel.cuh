// salp_pipe_kernel.cuh -- building blocks of the warp-specialised small-batch step kernel
// (salp_pipe4_kernel.cuh): named-barrier hand-off and the shared-memory row formats of the rings.
//
// The shape of the body and every coefficient derived from it depend on the action and the substep
// index only, never on the motion state; the kinematics depend on (v, w) only.  So the substep is cut
// along that feed-forward structure into instruction streams that run on different warps of one
// block (one block = 32 envs) and talk through shared-memory rings:
//   ring 1: ShapeFront (fp64 shape chain + backward differences, rounded to fp32)   8 floats / lane
//   ring 2: the rotational half of Coef32 (the fp32 coefficient set of a substep)   8 or 16 floats / lane
//   ring 3: (v, w) after the dynamics of a substep                                   6 floats / lane
// Slot = substep mod ring size, one 16-byte-aligned row per lane (conflict-free LDS.128 / STS.128).
// Hand-off is chunk-granular: bar.arrive by the side that is done with a chunk, bar.sync by the
// side that needs it, so nobody waits unless a neighbour has fallen a whole chunk behind.
#pragma once
#include "salp_env.cuh"

// named barriers (0 is __syncthreads)
__device__ __forceinline__ void pipe_bar_sync(int id, int threads = 64) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
// (no fence: a completed barrier orders the shared-memory accesses its participants made before
//  arriving -- the producer/consumer idiom of the PTX ISA's bar.arrive / bar.sync example; the
//  hand-offs are checked by the tag planes of SALP_STEP_CHECK_HANDOFF, salp_pipe4_kernel.cuh)
__device__ __forceinline__ void pipe_bar_arrive(int id, int threads = 64) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// Rings are planes of quads: row pointer = first quad of this lane, quads are 32 float4 apart.
__device__ __forceinline__ void front_store(const ShapeFront& f, float4* row) {
  row[0] = make_float4(f.dl, f.I_rate0, f.I_rate1, f.dV_dt);
  row[32] = make_float4(f.com, f.com_rate, f.com_acc, f.jet_on);
}
__device__ __forceinline__ void front_load(ShapeFront& f, const float4* row) {
  const float4 a = row[0], b = row[32];
  f.dl = a.x; f.I_rate0 = a.y; f.I_rate1 = a.z; f.dV_dt = a.w;
  f.com = b.x; f.com_rate = b.y; f.com_acc = b.z; f.jet_on = b.w;
}
// Rotational half of the coefficient set (make_coefs_R): two quads for the axisymmetric form, four
// otherwise.  (The translational half never leaves the dyn warp's registers.)
template <bool AXI>
__device__ __forceinline__ void coef_store_R(const Coef32& g, float4* row) {
  row[0] = make_float4(g.tj1, g.tj2, g.kqI[0], g.kqI[1]);
  row[32] = make_float4(g.klI[0], g.klI[1], g.JdI[1], g.AdI[1]);
  if (!AXI) {
    row[64] = make_float4(g.kqI[2], g.klI[2], g.JdI[0], g.JdI[2]);
    row[96] = make_float4(g.AdI[0], g.AdI[2], 0.f, 0.f);
  }
}
template <bool AXI>
__device__ __forceinline__ void coef_load_R(Coef32& g, const float4* row) {
  const float4 d = row[0], e = row[32];
  g.tj1 = d.x; g.tj2 = d.y; g.kqI[0] = d.z; g.kqI[1] = d.w;
  g.klI[0] = e.x; g.klI[1] = e.y; g.JdI[1] = e.z; g.AdI[1] = e.w;
  if (!AXI) {
    const float4 f = row[64], h = row[96];
    g.kqI[2] = f.x; g.klI[2] = f.y; g.JdI[0] = f.z; g.JdI[2] = f.w;
    g.AdI[0] = h.x; g.AdI[2] = h.y;
  }
}
