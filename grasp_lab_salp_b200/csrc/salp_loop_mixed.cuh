// salp_loop_mixed.cuh -- SALP_PRECISION_MIXED: the production substep loop.
//
// Same Robot.step() (robot.py:670-678) as salp_loop_f64.cuh, re-organised for the FP32 pipe:
//
//  * Motion state (v, w, a_prev, alpha_prev, roll, pitch) lives in fp32 registers for all K
//    substeps of the cycle.  The dynamics are dissipative (SURVEY.md hard part 4), so fp32
//    rounding does not amplify; agreement with the float64 reference is ~1e-6 relative per
//    env-step (tests/test_gpu_parity.py states 1e-5).
//  * Everything the reference *differences* -- water volume (jet speed, mass rate), inertia
//    (deformation torque), centre of mass (first and second backward difference, the second one
//    amplified by 1/dt^2 = 1e4) -- is evaluated in fp64 from dl = init_length - length, then
//    rounded to fp32 once.  The fp64 chain runs ONLY while the shape changes (refill ramp, jet)
//    plus two settle substeps; in the hold / coast / rest phases (most of a cycle) every finite
//    difference is exactly 0 in the reference too, and the ~30 geometry-derived coefficients
//    stay in registers (phase-specialised loop, SURVEY.md hard part 3).
//  * Integrals that grow over an episode (world position, body-frame position/angle integrals,
//    the three Euler angles) are two-level sums: an fp32 partial per 32-substep chunk, flushed
//    into an fp64 total.  sin/cos of each Euler angle is carried as a pair that is rotated by the
//    small per-substep increment (5-instruction Taylor kernels, no range reduction) and
//    re-anchored from the fp64 total at every flush, so it is valid for any angle.
//  * kin(k-1) and dyn(k) are software-pipelined into one basic block (two independent chains).
//  * The substep count K and the phase of every substep are decided exactly as the reference
//    does (float32/float64 comparison quirks of SURVEY.md hard part 2) from the table t_k of
//    k-fold repeated `cycle_time += 0.01` additions.
#pragma once
#include "salp_loop_f64.cuh"

#define SALP_MIXED_CHUNK 32

// ---- MUFU-based reciprocal / square root ----------------------------------------------------
// The bare MUFU results (rcp.approx / sqrt.approx: max relative error 2^-23 resp. ~1 ulp per the
// PTX ISA, i.e. at the level of one fp32 rounding) -- one instruction on the v -> |v| -> a -> v
// recurrence and in the coefficient set instead of seed + Newton step (6 dependent instructions).
SALP_HD float loop_rcp(float x) {
#ifdef __CUDA_ARCH__
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return 1.0f / x;
#endif
}
// bc2 = b^2 + c^2, the partial sum on the way (the fictitious force needs w1^2 + w2^2)
SALP_HD float fast_norm3(float a, float b, float c, float& bc2) {
  bc2 = fmaf(b, b, rn::fmul(c, c));
  const float s = fmaf(a, a, bc2);
#ifdef __CUDA_ARCH__
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(s));
  return r;
#else
  return sqrtf(s);
#endif
}

SALP_HD float fast_norm3(float a, float b, float c) {
  float bc2;
  return fast_norm3(a, b, c, bc2);
}

// np_sincosf without the separately-rounded steps: same Cody-Waite + minimax kernels (1 ulp),
// free to contract.  |x| <= 71476.  Used outside the substep loop (chunk anchors).
SALP_HD void sincos32(float x, float& sn, float& cs) {
  float q = rn::fsub(fmaf(x, 0x1.45f306p-1f, 0x1.8p+23f), 0x1.8p+23f);
  float r = fmaf(q, -0x1.921fb0p+0f, x);
  r = fmaf(q, -0x1.5110b4p-22f, r);
  r = fmaf(q, -0x1.846988p-48f, r);
  float r2 = r * r;
  float C = fmaf(fmaf(fmaf(fmaf(0x1.98e616p-16f, r2, -0x1.6c06dcp-10f), r2, 0x1.55553cp-5f), r2, -0.5f), r2, 1.0f);
  float S = fmaf(fmaf(fmaf(fmaf(0x1.7d3bbcp-19f, r2, -0x1.a06bbap-13f), r2, 0x1.11119ap-7f), r2, -0x1.555556p-3f) * r2, r, r);
  int k = (int)q;
  cs = (k & 1) ? S : C;
  sn = (k & 1) ? C : S;
  if ((k + 1) & 2) cs = -cs;
  if (k & 2) sn = -sn;
}
// |x| <= 0.55: Taylor to x^9 / x^8 (truncation < 2e-8 relative), branch-free, 10 instructions.
SALP_HD void sincos_small(float x, float& sn, float& cs) {
  float x2 = x * x;
  float ps = fmaf(fmaf(fmaf(2.7557319e-6f, x2, -1.9841270e-4f), x2, 8.3333333e-3f), x2, -1.6666667e-1f);
  sn = fmaf(x * x2, ps, x);
  cs = fmaf(fmaf(fmaf(fmaf(2.4801587e-5f, x2, -1.3888889e-3f), x2, 4.1666667e-2f), x2, -0.5f), x2, 1.0f);
}

// Host-derived constants of the mixed loop (computed once per launch from SalpParams, passed as a
// kernel argument so that no double->float conversion or constant folding is left in the loop).
struct SalpDerived {
  // fp64 shape chain
  double inv_dt, four_thirds_pi, skin3, c2, c1, c0, comA, comB, mtot0, m0, jet_gain;
  // fp32
  float dt, ratio_f, pi, end_aspect, inv_aspect_span, half_rho_neg, half_rho_pi, torque_ratio, arm0;
  float init_length_f, init_width_f, jet_gain_f, rho_f, m_base_f, four_thirds_pi_f, skin3_f, c2_f, c1_f, c0_f;   // fp32 shape (never differenced)
  float Ca[3], E[3], Cat[3], Car[3], CaD[3], CatF[3];
  float thi[3], tspan[3], rhi[3], rspan[3];
  // 1 if axes 1 and 2 carry the same coefficients (the default parameters do): the loop then
  // shares their coefficient entries and drops the terms whose coefficient is identically 0
  // (bit-identical results; see make_coefs / dyn_step)
  int axisym;
};

SALP_HD SalpDerived make_derived(const SalpParams& p) {
  SalpDerived k;
  k.inv_dt = 1.0 / p.dt;
  k.four_thirds_pi = (4.0 / 3.0) * M_PI;
  // geometry.py:137-141 literals
  const double mass_buoy = 0.195, skin_mass = 0.145, tube_mass = 0.414;
  const double tube_volume = 3.14159265358979 * ((0.058 / 2.0) * (0.058 / 2.0)) * 0.15;
  const double ntm = tube_mass - tube_volume * 1000.0;
  const double nm = p.nozzle_mass;
  k.skin3 = skin_mass / 3.0;
  // buoy*lh^2 + ntm*(lh-0.08)^2 + nm*(lh+0.025)^2 = c2 lh^2 + c1 lh + c0
  k.c2 = mass_buoy + ntm + nm;
  k.c1 = -0.16 * ntm + 0.05 * nm;
  k.c0 = 0.0064 * ntm + 0.000625 * nm;
  // geometry.py:187-203: water_mass * pos_water == -density * tube_volume * pos_tube exactly
  // (water_mass = density*V and wme - tv = 1000*V), so the numerator is linear in lh
  const double A_t = p.tube_mass - p.density * p.tube_volume;
  k.comA = A_t - nm + p.buoy_mass;                 // pos_tube = lh-0.08, pos_nozzle = 0.025-lh, pos_buoy = lh
  k.comB = -0.08 * A_t + 0.025 * nm;
  k.mtot0 = p.tube_mass + nm + p.buoy_mass + p.skin_mass;
  k.m0 = p.dry_mass + nm;
  // F_jet = -Cd * mass_rate * (dV/dt / A_nozzle) * dir,  mass_rate = rho dV/dt   (dynamics.py:88-101)
  k.jet_gain = -p.discharge_coefficient * p.density / p.nozzle_area;
  k.init_length_f = (float)p.init_length;
  k.init_width_f = (float)p.init_width;
  k.jet_gain_f = (float)k.jet_gain;
  k.rho_f = (float)p.density;
  k.m_base_f = (float)(p.dry_mass + nm - p.density * p.tube_volume);    // m = dry + nozzle + rho (V_ell - V_tube)
  k.four_thirds_pi_f = (float)k.four_thirds_pi;
  k.skin3_f = (float)k.skin3;
  k.c2_f = (float)k.c2;
  k.c1_f = (float)k.c1;
  k.c0_f = (float)k.c0;
  k.dt = (float)p.dt;
  k.ratio_f = (float)p.drag_force_ratio;
  k.pi = (float)M_PI;
  const double init_aspect = p.init_length / p.init_width;
  const double end_aspect = (p.init_length - p.max_contraction) / (p.max_contraction + p.init_width);
  k.end_aspect = (float)end_aspect;
  k.inv_aspect_span = (float)(1.0 / (init_aspect - end_aspect));
  k.half_rho_neg = (float)(-0.5 * p.density);
  k.half_rho_pi = (float)(-0.5 * p.density * M_PI);
  k.torque_ratio = (float)p.drag_torque_ratio;
  k.arm0 = -(float)(p.nozzle_length1 + p.nozzle_length2);
  for (int i = 0; i < 3; i++) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
    k.Ca[i] = (float)p.added_mass_force[i];
    k.E[i] = (float)(1.0 + p.added_mass_force[i]);
    k.Cat[i] = (float)p.added_mass_torque[i];
    k.Car[i] = (float)p.added_mass_rate_force[i];
    k.CaD[i] = (float)(p.added_mass_force[i2] - p.added_mass_force[i1]);       // (v x (Ca o v))_i = v_i1 v_i2 CaD_i
    k.CatF[i] = (float)(1.0 + p.added_mass_torque[i]);
    k.thi[i] = (float)p.trans_drag_range[2 * i + 1];
    k.tspan[i] = (float)(p.trans_drag_range[2 * i + 1] - p.trans_drag_range[2 * i]);
    k.rhi[i] = (float)p.rot_drag_range[2 * i + 1];
    k.rspan[i] = (float)(p.rot_drag_range[2 * i + 1] - p.rot_drag_range[2 * i]);
  }
  k.axisym = k.Ca[1] == k.Ca[2] && k.Cat[1] == k.Cat[2] && k.Car[0] == k.Car[1] && k.Car[1] == k.Car[2] &&
             k.thi[1] == k.thi[2] && k.tspan[1] == k.tspan[2] && k.rhi[1] == k.rhi[2] && k.rspan[1] == k.rspan[2];
  return k;
}

// fp32 coefficient set of one substep: everything the Newton/Euler equations need from the body
// shape, pre-divided by mass / inertia.  Loop-invariant while the shape is static.  With
//   M = m, Ca/Car/Cat the added-mass diagonals, E = 1 + Ca, J_i = I_i (1 + Cat_i):
//   a_i     = aj_i + v_i (kdm_i |v| + xc_i) - Ca_i a_prev,i - (w x (E o v))_i + fict_i,  xc_i = kdm_i ratio - mrm_i
//   alpha_i = tj_i + w_i (kqI_i |w| + klI_i) - Cat_i alpha_prev,i - w_i1 w_i2 JdI_i - v_i1 v_i2 AdI_i
// which is robot.py:789-851 / dynamics.py:6-174 with the common factors 1/m, 1/I_i cancelled
// (Coriolis and added-mass cross products merged; the two cross products of a vector with its own
// diagonal scaling collapse to one product per component).
struct Coef32 {
  float aj[3];        // F_jet / m                                       (robot.py:937-951)
  float kdm[3];       // -rho/2 area_i Ct_i / m                          (dynamics.py:111-116)
  float xc[3];        // kdm_i ratio - mass_rate Car_i / m                (dynamics.py:111-116, :139)
  float com, com_rate2, com_acc;      // centre of mass, 2 x its rate, its acceleration     robot.py:898-922
  float tj1, tj2;     // (arm x F_jet)_i / I_i                           (robot.py:931-935)
  float kqI[3];       // -rho/2 Cr_i area_i dims_i / I_i                 (dynamics.py:120-128)
  float klI[3];       // (ratio -rho/2 Cr_i area_i width - I_rate_i) / I_i   (+ deform torque, :172-174)
  float JdI[3];       // (J_i2 - J_i1) / I_i
  float AdI[3];       // m (Ca_i2 - Ca_i1) / I_i
  float inv_m, inv_Iz;  // only read by the disturbance variant (force / torque noise enter as F/m, T/I)
};

// fp64 side of the shape: the quantities that are differenced.
struct Shape64 {
  double V;           // water volume (ellipsoid - tube), robot.py:1055-1056
  double I0, I1;      // inertia diagonal
  double com;         // centre of mass x
  double com_rate;
};

// fp64 shape chain at half-length lh, half-width wh: 19 flop + 1 division.
SALP_HD void shape64_at(const SalpParams& p, const SalpDerived& k, double lh, double wh, double& V,
                        double& I0, double& I1, double& com, double& wm) {
  double wh2 = wh * wh, lh2 = lh * lh;
  double Ve = k.four_thirds_pi * lh * wh2;
  V = Ve - p.tube_volume;
  wm = p.density * V;
  double sw = k.skin3 + 200.0 * Ve;               // skin/3 + 0.2 * 1000 * V_ellipsoid
  I0 = sw * (wh2 + wh2);
  I1 = (k.c2 * lh2 + k.c1 * lh + k.c0) + sw * (lh2 + wh2);
  com = (k.comA * lh + k.comB) / (k.mtot0 + wm);
}
// The same chain inside the substep loop.  The division by the total mass D is a hardware
// reciprocal seed (rcp.approx.ftz.f64, ~20 good bits) refined by two Newton steps (2^-20 -> 2^-40
// -> below fp64 rounding): 5 dependent operations instead of a ~200-cycle DDIV, a pure function of
// D (an unchanged shape reproduces every bit: redundant updates stay exact no-ops), and no value
// carried from the previous update, so consecutive updates are independent chains.
SALP_HD double rcp64_seed(double x) {
#ifdef __CUDA_ARCH__
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  return r;
#else
  return (double)(1.0f / (float)x);
#endif
}
SALP_HD void shape64_step(const SalpParams& p, const SalpDerived& k, double lh, double wh,
                          double& V, double& I0, double& I1, double& com) {
  double wh2 = wh * wh, lh2 = lh * lh;
  double Ve = k.four_thirds_pi * lh * wh2;
  V = Ve - p.tube_volume;
  double D = k.mtot0 + p.density * V;
  double sw = k.skin3 + 200.0 * Ve;
  I0 = sw * (wh2 + wh2);
  I1 = (k.c2 * lh2 + k.c1 * lh + k.c0) + sw * (lh2 + wh2);
  double r = rcp64_seed(D);
  r = r * (2.0 - D * r);
  r = r * (2.0 - D * r);
  com = (k.comA * lh + k.comB) * r;
}

// All fp32 coefficients of the coming substep.  Mass, inertia, areas and drag coefficients are
// evaluated in fp32 from the half-length / half-width (they are never differenced); the
// differenced quantities arrive from the fp64 chain already rounded.
// AXI (SalpDerived.axisym): entries [2] of kdm / xc / kqI / klI equal entries [1],
// JdI[0] = AdI[0] = 0, JdI[2] = -JdI[1], AdI[2] = -AdI[1] -- they are neither computed nor read.
//
// The set comes in two halves that the five-warp pipeline kernel computes on different warps:
//   make_coefs_T  translational: jet acceleration, drag / mass-rate terms, centre-of-mass terms
//   make_coefs_R  rotational:    jet torque, rotational drag, deformation, inertia ratios
// Both start from the same handful of shape quantities (CoefBase, ~20 instructions, computed by
// each half).  EVERY operation is explicitly rounded (rn:: / fmaf), so the halves give the same
// bits wherever they are inlined -- the fused kernel's make_coefs is just base + T + R.
struct CoefBase {
  float wh2, lh2, lw, m, sw, nr, P0, P1, f;
};
SALP_HD void coef_base(const SalpDerived& k, bool jet_on, float lh, float wh, float dV_dt, CoefBase& c) {
  c.wh2 = rn::fmul(wh, wh);
  c.lh2 = rn::fmul(lh, lh);
  c.lw = rn::fmul(lh, wh);
  const float Ve = rn::fmul(k.four_thirds_pi_f, rn::fmul(c.lw, wh));   // geometry.py:79-81
  c.m = fmaf(k.rho_f, Ve, k.m_base_f);                                 // robot.py:1055-1063
  c.sw = fmaf(200.0f, Ve, k.skin3_f);                                  // geometry.py:134-183
  // aspect-ratio interpolation of the drag coefficients (geometry.py:105-123)
  const float nr = rn::fmul(fmaf(lh, loop_rcp(wh), -k.end_aspect), k.inv_aspect_span);
  c.nr = fminf(fmaxf(nr, 0.0f), 1.0f);
  // -rho/2 * area_i (geometry.py:68-75: areas pi wh^2, pi lh wh, pi lh wh)
  c.P0 = rn::fmul(k.half_rho_pi, c.wh2);
  c.P1 = rn::fmul(k.half_rho_pi, c.lw);
  c.f = jet_on ? rn::fmul(rn::fmul(k.jet_gain_f, dV_dt), dV_dt) : 0.0f;   // F_jet = -Cd rho (dV/dt)^2 / A_nozzle
}
template <bool AXI = false>
SALP_HD void make_coefs_T(const SalpDerived& k, const CoefBase& c, const float dir[3], float dV_dt, float com,
                          float com_rate, float com_acc, Coef32& g) {
  const float inv_m = loop_rcp(c.m);
  const float Q0 = rn::fmul(c.P0, inv_m), Q1 = rn::fmul(c.P1, inv_m);
  g.kdm[0] = rn::fmul(Q0, fmaf(-c.nr, k.tspan[0], k.thi[0]));
  g.kdm[1] = rn::fmul(Q1, fmaf(-c.nr, k.tspan[1], k.thi[1]));
  const float mr = rn::fmul(rn::fmul(k.rho_f, dV_dt), inv_m);         // mass_rate / m   (geometry.py:98-101)
  // v_i (kdm_i (|v| + ratio) - mr Car_i) = v_i (kdm_i |v| + xc_i)
  g.xc[0] = fmaf(g.kdm[0], k.ratio_f, -rn::fmul(mr, k.Car[0]));
  g.xc[1] = fmaf(g.kdm[1], k.ratio_f, -rn::fmul(mr, k.Car[1]));
  if (!AXI) {
    g.kdm[2] = rn::fmul(Q1, fmaf(-c.nr, k.tspan[2], k.thi[2]));
    g.xc[2] = fmaf(g.kdm[2], k.ratio_f, -rn::fmul(mr, k.Car[2]));
  }
  const float fm = rn::fmul(c.f, inv_m);
  g.aj[0] = rn::fmul(dir[0], fm); g.aj[1] = rn::fmul(dir[1], fm); g.aj[2] = rn::fmul(dir[2], fm);
  g.com = com;
  g.com_rate2 = rn::fadd(com_rate, com_rate);
  g.com_acc = com_acc;
  g.inv_m = inv_m;
}
template <bool AXI = false>
SALP_HD void make_coefs_R(const SalpDerived& k, const CoefBase& c, const float dir[3], float lh, float wh,
                          float I_rate0, float I_rate1, Coef32& g) {
  const float swh = rn::fmul(c.sw, c.wh2);
  const float I0 = rn::fadd(swh, swh);
  const float I1 = fmaf(c.sw, rn::fadd(c.lh2, c.wh2), fmaf(k.c2_f, c.lh2, fmaf(k.c1_f, lh, k.c0_f)));
  const float inv_I0 = loop_rcp(I0), inv_I1 = loop_rcp(I1);
  const float kr0 = rn::fmul(c.P0, fmaf(-c.nr, k.rspan[0], k.rhi[0]));
  const float kr1 = rn::fmul(c.P1, fmaf(-c.nr, k.rspan[1], k.rhi[1]));
  // drag torque: dims = (width^3, length^3, length^3) = 8 (wh^3, lh^3, lh^3)
  const float E0 = rn::fmul(rn::fmul(c.wh2, wh), rn::fmul(8.0f, inv_I0));
  const float E1 = rn::fmul(rn::fmul(c.lh2, lh), rn::fmul(8.0f, inv_I1));
  g.kqI[0] = rn::fmul(kr0, E0); g.kqI[1] = rn::fmul(kr1, E1);
  const float tw = rn::fmul(k.torque_ratio, rn::fadd(wh, wh));
  g.klI[0] = rn::fmul(fmaf(kr0, tw, -I_rate0), inv_I0);
  g.klI[1] = rn::fmul(fmaf(kr1, tw, -I_rate1), inv_I1);
  if (!AXI) {
    const float kr2 = rn::fmul(c.P1, fmaf(-c.nr, k.rspan[2], k.rhi[2]));
    g.kqI[2] = rn::fmul(kr2, E1);
    g.klI[2] = rn::fmul(fmaf(kr2, tw, -I_rate1), inv_I1);
  }
  // (J_i2 - J_i1) / I_i with J = I o (1 + Cat), I = (I0, I1, I1)
  const float r01 = rn::fmul(I0, inv_I1);
  g.JdI[1] = fmaf(r01, k.CatF[0], -k.CatF[2]);
  const float mI1 = rn::fmul(c.m, inv_I1);
  g.AdI[1] = rn::fmul(mI1, k.CaD[1]);
  if (!AXI) {
    g.JdI[0] = rn::fmul(rn::fmul(I1, inv_I0), rn::fsub(k.CatF[2], k.CatF[1]));
    g.JdI[2] = fmaf(-r01, k.CatF[0], k.CatF[1]);
    g.AdI[0] = rn::fmul(rn::fmul(c.m, inv_I0), k.CaD[0]);
    g.AdI[2] = rn::fmul(mI1, k.CaD[2]);
  }
  const float afI = rn::fmul(rn::fsub(k.arm0, lh), rn::fmul(c.f, inv_I1));   // arm = (arm0 - lh, 0, 0)   robot.py:931-935
  g.tj1 = rn::fmul(-afI, dir[2]);
  g.tj2 = rn::fmul(afI, dir[1]);
  g.inv_Iz = inv_I1;
}
template <bool AXI = false>
SALP_HD void make_coefs(const SalpDerived& k, const float dir[3], bool jet_on, float lh, float wh,
                        float I_rate0, float I_rate1, float dV_dt, float com, float com_rate, float com_acc,
                        Coef32& g) {
  CoefBase c;
  coef_base(k, jet_on, lh, wh, dV_dt, c);
  make_coefs_T<AXI>(k, c, dir, dV_dt, com, com_rate, com_acc, g);
  make_coefs_R<AXI>(k, c, dir, lh, wh, I_rate0, I_rate1, g);
}

// fp32 register state of one env inside the substep loop
struct Motion32 {
  float v0, v1, v2, w0, w1, w2;          // velocity, angular_velocity (body frame)
  float ac0, ac1, ac2, al0, al1, al2;    // previous substep's accelerations (robot.py:806, 992, 1005)
  // Euler angles: fp64 totals live in Body64.eul; here the increments since the last flush and the
  // (sin, cos) pairs, each = fp64-anchored value rotated by the fp32 increments of this chunk
  // (sin/cos of roll and pitch are carried: the Euler-rate matrix uses the OLD angles)
  float phi_lo, theta_lo, psi_lo;
  float sph, cph, sth, cth, sps, cps;
  float pw0, pw1, pw2;                   // chunk partial sums: position_world, position, angle
  float pos0, pos1, pos2, ang0, ang1, ang2;
  float vw0, vw1;                        // velocity_world[0:2] of the latest kinematic update
};

// _newton_equations + _euler_equations + the velocity half of _update_motion_states
// (robot.py:789-862) in the pre-divided form documented at Coef32.  ~85 FP32 instructions.
// OUDisturbance.sample() for the three live components (robot.py:236-242): Euler-Maruyama step
// x += theta (0 - x) dt + sigma sqrt(dt) N(0, 1), theta = 2, sigma = 0.05 (force) / 0.01 (torque)
SALP_HD void ou_step(const SalpDerived& dv, RandCtx& rc, int k) {
  uint32_t r[4];
  rand_block(rc.seed, rc.gid, rc.episode, rc.cycle, SALP_RNG_OU + (uint32_t)k, r);
  const float two_pi = 6.2831853f;
#ifdef __CUDA_ARCH__
  float m0 = sqrtf(-2.0f * __logf(u01(r[0]))), m1 = sqrtf(-2.0f * __logf(u01(r[2])));
  float n0 = m0 * __cosf(two_pi * u01(r[1])), n1 = m0 * __sinf(two_pi * u01(r[1])), n2 = m1 * __cosf(two_pi * u01(r[3]));
#else
  float m0 = sqrtf(-2.0f * logf(u01(r[0]))), m1 = sqrtf(-2.0f * logf(u01(r[2])));
  float n0 = m0 * cosf(two_pi * u01(r[1])), n1 = m0 * sinf(two_pi * u01(r[1])), n2 = m1 * cosf(two_pi * u01(r[3]));
#endif
  const float decay = 1.0f - 2.0f * dv.dt, sq = sqrtf(dv.dt);
  rc.ou_fx = rc.ou_fx * decay + 0.05f * sq * n0;
  rc.ou_fy = rc.ou_fy * decay + 0.05f * sq * n1;
  rc.ou_tz = rc.ou_tz * decay + 0.01f * sq * n2;
}

// STATIC: the shape has stopped moving (part B of the loop, after the warp-uniform end of every
// lane's update window + its two settle substeps): com_rate and com_acc are exactly 0 there.
template <bool NOISE = false, bool STATIC = false, bool AXI = false>
SALP_HD void dyn_step(const SalpDerived& dv, const Coef32& g, Motion32& s, RandCtx* rc = nullptr, int k = 0) {
  const float v0 = s.v0, v1 = s.v1, v2 = s.v2, w0 = s.w0, w1 = s.w1, w2 = s.w2;
  // The recurrence v -> |v| -> a -> v is the critical path of the whole kernel when the GPU is not
  // full: |v| and |w| are one MUFU.SQRT each, and they enter LAST -- everything that does not depend
  // on them (added mass, Coriolis, fictitious forces, the jet) is summed first, in the shadow of
  // the square roots:   a_i = base_i + v_i X_i,  X_i = kdm_i |v| + xc_i      (|v| v + ratio v = v (|v| + ratio))
  const float nv = fast_norm3(v0, v1, v2);
  float w12;                                                    // w1^2 + w2^2
  const float wn = fast_norm3(w0, w1, w2, w12);
  // Every product and sum below is an explicitly rounded operation (fmaf / rn::), so the compiler
  // has no freedom in how it contracts a*b+c: the general form, the axisymmetric form and the
  // static-shape form round identically wherever they compute the same quantity, in every kernel.
  const float ev0 = rn::fmul(dv.E[0], v0), ev1 = rn::fmul(dv.E[1], v1), ev2 = rn::fmul(dv.E[2], v2);
  const float p20 = rn::fmul(w2, w0), p01 = rn::fmul(w0, w1);
  const float q20 = rn::fmul(v2, v0), q01 = rn::fmul(v0, v1);
  // a_i - v_i X_i = aj_i - Ca_i a_prev,i - (w x (E o v))_i + fict_i
  float ba0 = fmaf(w2, ev1, fmaf(-w1, ev2, fmaf(-dv.Ca[0], s.ac0, g.aj[0])));
  float ba1 = fmaf(w0, ev2, fmaf(-w2, ev0, fmaf(-dv.Ca[1], s.ac1, g.aj[1])));
  float ba2 = fmaf(w1, ev0, fmaf(-w0, ev1, fmaf(-dv.Ca[2], s.ac2, g.aj[2])));
  // fictitious forces of the moving centre of mass c = (com, 0, 0) (robot.py:806-810):
  //   -(alpha x c) - w x (w x c) - 2 w x c' - c''  with the products of w shared with the Euler equations
  const float t1 = rn::fadd(s.al2, p01), t2 = rn::fsub(p20, s.al1);
  if (!STATIC) {      // (com_rate = com_acc = 0 exactly while the shape is static: these three are then exact no-ops)
    ba0 = rn::fadd(ba0, g.com_acc);
    ba1 = fmaf(w2, g.com_rate2, ba1);
    ba2 = fmaf(-w1, g.com_rate2, ba2);
  }
  ba0 = fmaf(-g.com, w12, ba0);
  ba1 = fmaf(g.com, t1, ba1);
  ba2 = fmaf(g.com, t2, ba2);
  // alpha_i - w_i Y_i = tj_i - Cat_i alpha_prev,i - w_i1 w_i2 JdI_i - v_i1 v_i2 AdI_i
  // AXI: kdm[2] = kdm[1], xc[2] = xc[1], kqI[2] = kqI[1], klI[2] = klI[1], JdI[0] = AdI[0] = 0,
  // JdI[2] = -JdI[1], AdI[2] = -AdI[1]
  float bl0 = rn::fmul(-dv.Cat[0], s.al0);
  float bl1 = fmaf(-q20, g.AdI[1], fmaf(-p20, g.JdI[1], fmaf(-dv.Cat[1], s.al1, g.tj1)));
  float bl2 = fmaf(-dv.Cat[2], s.al2, g.tj2);
  if (AXI) {
    bl2 = fmaf(q01, g.AdI[1], fmaf(p01, g.JdI[1], bl2));
  } else {
    const float p12 = rn::fmul(w1, w2), q12 = rn::fmul(v1, v2);
    bl0 = fmaf(-q12, g.AdI[0], fmaf(-p12, g.JdI[0], bl0));
    bl2 = fmaf(-q01, g.AdI[2], fmaf(-p01, g.JdI[2], bl2));
  }
  if (NOISE) {        // force_noise / torque_noise join the sums of _newton_equations / _euler_equations
    ou_step(dv, *rc, k);
    ba0 = fmaf(rc->ou_fx, g.inv_m, ba0);
    ba1 = fmaf(rc->ou_fy, g.inv_m, ba1);
    bl2 = fmaf(rc->ou_tz, g.inv_Iz, bl2);
  }
  const float X0 = fmaf(g.kdm[0], nv, g.xc[0]);
  const float X1 = fmaf(g.kdm[1], nv, g.xc[1]);
  const float X2 = AXI ? X1 : fmaf(g.kdm[2], nv, g.xc[2]);
  const float Y0 = fmaf(g.kqI[0], wn, g.klI[0]);
  const float Y1 = fmaf(g.kqI[1], wn, g.klI[1]);
  const float Y2 = AXI ? Y1 : fmaf(g.kqI[2], wn, g.klI[2]);
  const float na0 = fmaf(v0, X0, ba0), na1 = fmaf(v1, X1, ba1), na2 = fmaf(v2, X2, ba2);
  const float nl0 = fmaf(w0, Y0, bl0), nl1 = fmaf(w1, Y1, bl1), nl2 = fmaf(w2, Y2, bl2);
  s.ac0 = na0; s.ac1 = na1; s.ac2 = na2;
  s.al0 = nl0; s.al1 = nl1; s.al2 = nl2;
  s.v0 = fmaf(na0, dv.dt, v0); s.v1 = fmaf(na1, dv.dt, v1); s.v2 = fmaf(na2, dv.dt, v2);
  s.w0 = fmaf(nl0, dv.dt, w0); s.w1 = fmaf(nl1, dv.dt, w1); s.w2 = fmaf(nl2, dv.dt, w2);
}

// The kinematic half of _update_motion_states (robot.py:864-875): Euler-angle rates at the old
// roll/pitch (dynamics.py:21-31), new angles, body->world rotation Rz Ry Rx (dynamics.py:35-58) by
// successive elementary rotations, the three position/angle integrals.  ~70 FP32 instructions.
// v, w are the velocities AFTER dyn_step of the same substep.
// (sin, cos)(x) -> (sin, cos)(x + d).  The Taylor kernels are valid for |d| <= 0.55 (a substep moves an
// Euler angle by ~1e-2 rad); a tumbling body passing the pitch = +-pi/2 singularity of the Euler-rate
// matrix (long random episodes do, in the reference too) can ask for more for a substep or two:
// the increment is clamped (kin_step) so that the pair stays a unit vector; the fp32 increments
// and the fp64 angle totals accumulate the same clamped values, so the pair and the angle agree
// and the pair is re-anchored from the total at the next flush.
// (explicitly rounded operations, like dyn_step: identical bits in every instantiation and kernel)
SALP_HD void rotate_small(float d, float& sn, float& cs) {
  // per-substep increments are ~1e-2 rad: sin d = d - d^3/6 + d^5/120, cos d = 1 - d^2/2 + d^4/24
  // (truncation < 1e-10 there, 2e-7 at the clamp; dropping the d^5 term was measured to triple the
  //  error of the violent cycles next to the integrator's stability limit)
  const float d2 = rn::fmul(d, d);
  const float sd_ = rn::fmul(d, fmaf(d2, fmaf(d2, 8.3333333e-3f, -1.6666667e-1f), 1.0f));
  const float cd_ = fmaf(d2, fmaf(d2, 4.1666667e-2f, -0.5f), 1.0f);
  const float ns = fmaf(sn, cd_, rn::fmul(cs, sd_));
  cs = fmaf(cs, cd_, -rn::fmul(sn, sd_));
  sn = ns;
}
SALP_HD float clamp_increment(float d) { return fminf(fmaxf(d, -0.25f), 0.25f); }
// kin_step = kin_world (Euler angles, world position) + kin_body (the body-frame integrals
// `position`, `angle` of robot.py:872-875).  The two halves share no state, so the four-warp
// pipeline kernel runs them on different warps (kin_body rides with the dynamics, which owns v, w).
SALP_HD void kin_world(const SalpDerived& dv, Motion32& s) {
  const float dt = dv.dt;
  const float v0 = s.v0, v1 = s.v1, v2 = s.v2, w0 = s.w0, w1 = s.w1, w2 = s.w2;
  const float rcth = loop_rcp(s.cth);          // (a pure function of cos(pitch): nothing carried, it changes sign when the body tumbles)
  const float q = fmaf(s.sph, w1, rn::fmul(s.cph, w2));
  // Euler rates (dynamics.py:21-31): psi' = q / cos(theta), phi' = w0 + sin(theta) psi', theta' = cph w1 - sph w2.
  // Only the 1 / cos(theta) terms can ask for more than the Taylor kernels take (see above): the
  // yaw increment is clamped, roll inherits the bound through sin(theta) psi', pitch is clamped too.
  const float dpsi = clamp_increment(rn::fmul(rn::fmul(q, rcth), dt));
  const float dphi = fmaf(s.sth, dpsi, rn::fmul(w0, dt));
  const float dtheta = clamp_increment(rn::fmul(fmaf(s.cph, w1, -rn::fmul(s.sph, w2)), dt));
  s.phi_lo = rn::fadd(s.phi_lo, dphi);
  s.theta_lo = rn::fadd(s.theta_lo, dtheta);
  s.psi_lo = rn::fadd(s.psi_lo, dpsi);
  rotate_small(dphi, s.sph, s.cph);
  rotate_small(dtheta, s.sth, s.cth);
  rotate_small(dpsi, s.sps, s.cps);
  const float u1 = fmaf(s.cph, v1, -rn::fmul(s.sph, v2)), u2 = fmaf(s.sph, v1, rn::fmul(s.cph, v2));      // Rx
  const float r0 = fmaf(s.cth, v0, rn::fmul(s.sth, u2)), vw2 = fmaf(s.cth, u2, -rn::fmul(s.sth, v0));     // Ry
  s.vw0 = fmaf(s.cps, r0, -rn::fmul(s.sps, u1));                                                          // Rz
  s.vw1 = fmaf(s.sps, r0, rn::fmul(s.cps, u1));
  s.pw0 = fmaf(s.vw0, dt, s.pw0); s.pw1 = fmaf(s.vw1, dt, s.pw1); s.pw2 = fmaf(vw2, dt, s.pw2);
}
SALP_HD void kin_body(const SalpDerived& dv, Motion32& s) {
  const float dt = dv.dt;
  s.pos0 = fmaf(s.v0, dt, s.pos0); s.pos1 = fmaf(s.v1, dt, s.pos1); s.pos2 = fmaf(s.v2, dt, s.pos2);
  s.ang0 = fmaf(s.w0, dt, s.ang0); s.ang1 = fmaf(s.w1, dt, s.ang1); s.ang2 = fmaf(s.w2, dt, s.ang2);
}
SALP_HD void kin_step(const SalpDerived& dv, Motion32& s) {
  kin_world(dv, s);
  kin_body(dv, s);
}

// (sin, cos) of an fp64 angle total, rounded to fp32: the chunk anchor.  The fp64 argument is split
// into hi = (float)x and lo = (float)(x - hi); sincos32(hi) is a 1-ulp fp32 evaluation and lo
// (<= half an ulp of hi) enters by the first-order angle-addition terms -- no fp64 sincos.
SALP_HD void anchor_sincos(double x, float& sn, float& cs) {
  const float hi = (float)x;
  const float lo = (float)(x - (double)hi);
  float s, c;
  sincos32(hi, s, c);
  sn = fmaf(c, lo, s);
  cs = fmaf(-s, lo, c);
}

// fold the fp32 chunk partials into the fp64 totals and re-anchor the three (sin, cos) pairs
// (flush_world / flush_body: the halves that belong to kin_world / kin_body)
SALP_HD void flush_world(Body64& b, Motion32& s) {
  b.pw[0] += (double)s.pw0; b.pw[1] += (double)s.pw1; b.pw[2] += (double)s.pw2;
  b.eul[0] += (double)s.phi_lo; b.eul[1] += (double)s.theta_lo; b.eul[2] += (double)s.psi_lo;
  anchor_sincos(b.eul[0], s.sph, s.cph);
  anchor_sincos(b.eul[1], s.sth, s.cth);
  anchor_sincos(b.eul[2], s.sps, s.cps);
  s.phi_lo = s.theta_lo = s.psi_lo = 0.f;
  s.pw0 = s.pw1 = s.pw2 = 0.f;
}
SALP_HD void flush_body(Body64& b, Motion32& s) {
  b.pos[0] += (double)s.pos0; b.pos[1] += (double)s.pos1; b.pos[2] += (double)s.pos2;
  b.ang[0] += (double)s.ang0; b.ang[1] += (double)s.ang1; b.ang[2] += (double)s.ang2;
  s.pos0 = s.pos1 = s.pos2 = 0.f;
  s.ang0 = s.ang1 = s.ang2 = 0.f;
}
SALP_HD void flush_chunk(Body64& b, Motion32& s) {
  flush_world(b, s);
  flush_body(b, s);
}

// Shape bookkeeping in fp64 (everything the reference differences) + the fp32 coefficient set.
struct ShapeTrack {
  Shape64 s;
  double prev_com_rate, com_acc, prevV, I0_prev_used, I1_prev_used, dl;
  int last_update;
};

// update_state + update_properties after substep j-1 (robot.py:640-668), only called while the
// shape moves or its backward differences have not been flushed yet.
template <bool AXI = false>
SALP_HD void shape_update_at(const SalpParams& p, const SalpDerived& dv, const CyclePlan& c, double t,
                             const float dir[3], int j, int k_T0, int k_jet, ShapeTrack& st, Coef32& g);
template <bool AXI = false>
SALP_HD void shape_update(const SalpParams& p, const SalpDerived& dv, const CyclePlan& c, const double* time_table,
                          const float dir[3], int j, int k_T0, int k_jet, ShapeTrack& st, Coef32& g) {
  shape_update_at<AXI>(p, dv, c, time_table[j], dir, j, k_T0, k_jet, st, g);
}
// The update in two stages, so that the pipeline kernel can run them on different warps:
//   shape_front : fp64 -- shape at t_j, the differenced quantities, rounded to fp32 once  (ShapeFront)
//   make_coefs  : fp32 -- the coefficient set from a ShapeFront (stateless)
struct ShapeFront {
  float dl;                       // init_length - length
  float I_rate0, I_rate1;         // (I - I_prev) / dt                     robot.py:887-896
  float dV_dt;                    // (V - V_prev) / dt
  float com, com_rate, com_acc;   // robot.py:898-922
  float jet_on;                   // 1 while the previous substep's state was JET (robot.py:937-951)
};
// t = t_j, the j-fold repeated `cycle_time += dt` (callers either read the table or carry the sum)
SALP_HD void shape_front(const SalpParams& p, const SalpDerived& dv, const CyclePlan& c, double t, int j, int k_T0,
                         int k_jet, ShapeTrack& st, ShapeFront& f) {
  const int phase = j < k_T0 ? 0 : (j < k_jet ? 1 : 2);
  st.dl = shape_delta(phase, t, c.refill, c.T0, (double)c.contraction32, c.contract_rate, c.release_rate);
  double lh = 0.5 * (p.init_length - st.dl), wh = 0.5 * (p.init_width + st.dl);
  double V, I0n, I1n, com;
  shape64_step(p, dv, lh, wh, V, I0n, I1n, com);
  double dV_dt = (V - st.s.V) * dv.inv_dt;
  double com_rate = (com - st.s.com) * dv.inv_dt;                // robot.py:901-910
  st.com_acc = (com_rate - st.prev_com_rate) * dv.inv_dt;        // robot.py:912-922
  st.prev_com_rate = com_rate;
  f.dl = (float)st.dl;
  f.I_rate0 = (float)((I0n - st.s.I0) * dv.inv_dt);
  f.I_rate1 = (float)((I1n - st.s.I1) * dv.inv_dt);
  f.dV_dt = (float)dV_dt;
  f.com = (float)com;
  f.com_rate = (float)com_rate;
  f.com_acc = (float)st.com_acc;
  f.jet_on = phase == 1 ? 1.0f : 0.0f;
  st.prevV = st.s.V;
  st.I0_prev_used = st.s.I0;
  st.I1_prev_used = st.s.I1;
  st.s.V = V; st.s.I0 = I0n; st.s.I1 = I1n; st.s.com = com; st.s.com_rate = com_rate;
  st.last_update = j;
}
SALP_HD float front_lh(const SalpDerived& dv, const ShapeFront& f) { return rn::fmul(0.5f, rn::fsub(dv.init_length_f, f.dl)); }
SALP_HD float front_wh(const SalpDerived& dv, const ShapeFront& f) { return rn::fmul(0.5f, rn::fadd(dv.init_width_f, f.dl)); }
template <bool AXI = false>
SALP_HD void make_coefs(const SalpDerived& dv, const float dir[3], const ShapeFront& f, Coef32& g) {
  make_coefs<AXI>(dv, dir, f.jet_on != 0.0f, front_lh(dv, f), front_wh(dv, f),
             f.I_rate0, f.I_rate1, f.dV_dt, f.com, f.com_rate, f.com_acc, g);
}
// the two halves from a ShapeFront (five-warp pipeline kernel)
template <bool AXI = false>
SALP_HD void make_coefs_T(const SalpDerived& dv, const float dir[3], const ShapeFront& f, Coef32& g) {
  CoefBase c;
  coef_base(dv, f.jet_on != 0.0f, front_lh(dv, f), front_wh(dv, f), f.dV_dt, c);
  make_coefs_T<AXI>(dv, c, dir, f.dV_dt, f.com, f.com_rate, f.com_acc, g);
}
template <bool AXI = false>
SALP_HD void make_coefs_R(const SalpDerived& dv, const float dir[3], const ShapeFront& f, Coef32& g) {
  CoefBase c;
  const float lh = front_lh(dv, f), wh = front_wh(dv, f);
  coef_base(dv, f.jet_on != 0.0f, lh, wh, f.dV_dt, c);
  make_coefs_R<AXI>(dv, c, dir, lh, wh, f.I_rate0, f.I_rate1, g);
}
template <bool AXI>
SALP_HD void shape_update_at(const SalpParams& p, const SalpDerived& dv, const CyclePlan& c, double t,
                             const float dir[3], int j, int k_T0, int k_jet, ShapeTrack& st, Coef32& g) {
  ShapeFront f;
  shape_front(p, dv, c, t, j, k_T0, k_jet, st, f);
  make_coefs<AXI>(dv, dir, f, g);
}

// ---- building blocks shared by the fused loop below and the pipeline kernel ------------------
// shape-derived state of the first substep (coefficient set g_0) from the carried columns
template <bool AXI = false>
SALP_HD void mixed_init_shape(const SalpParams& p, const SalpDerived& dv, const Body64& b, const float dir[3],
                              ShapeTrack& st, Coef32& g) {
  st.prev_com_rate = b.prev_com_rate;
  st.com_acc = b.com_acc;
  st.prevV = b.prev_volume;
  st.dl = p.init_length - b.length;
  st.last_update = 0;
  double lh = 0.5 * b.length, wh = 0.5 * b.width, wm, com_now;
  shape64_at(p, dv, lh, wh, st.s.V, st.s.I0, st.s.I1, com_now, wm);
  double dV_dt = (st.s.V - b.prev_volume) * dv.inv_dt;
  // the carried centre of mass may be stale w.r.t. length/width (Robot.reset quirk, robot.py:478)
  st.s.com = b.com;
  st.s.com_rate = b.com_rate;
  make_coefs<AXI>(dv, dir, b.phase == 1, (float)lh, (float)wh,
             (float)((st.s.I0 - b.prevI[0]) * dv.inv_dt), (float)((st.s.I1 - b.prevI[1]) * dv.inv_dt),
             (float)dV_dt, (float)b.com, (float)b.com_rate, (float)b.com_acc, g);
  st.I0_prev_used = st.s.I0;
  st.I1_prev_used = st.s.I1;
}
SALP_HD void mixed_init_dyn(const Body64& b, Motion32& s) {
  s.v0 = (float)b.v[0]; s.v1 = (float)b.v[1]; s.v2 = (float)b.v[2];
  s.w0 = (float)b.w[0]; s.w1 = (float)b.w[1]; s.w2 = (float)b.w[2];
  s.ac0 = (float)b.acc[0]; s.ac1 = (float)b.acc[1]; s.ac2 = (float)b.acc[2];
  s.al0 = (float)b.alp[0]; s.al1 = (float)b.alp[1]; s.al2 = (float)b.alp[2];
}
SALP_HD void mixed_init_kin(const Body64& b, Motion32& s) {
  anchor_sincos(b.eul[0], s.sph, s.cph);
  anchor_sincos(b.eul[1], s.sth, s.cth);
  anchor_sincos(b.eul[2], s.sps, s.cps);
  s.phi_lo = s.theta_lo = s.psi_lo = 0.f;
  s.pw0 = s.pw1 = s.pw2 = 0.f;
  s.pos0 = s.pos1 = s.pos2 = 0.f; s.ang0 = s.ang1 = s.ang2 = 0.f;
  s.vw0 = s.vw1 = 0.f;
}
// back to the carried fp64 columns
SALP_HD void mixed_finish_shape(const SalpParams& p, ShapeTrack& st, int K, Body64& b) {
  if (st.last_update != K) {     // static tail: update_properties re-assigned the same shape (robot.py:651-668)
    st.prevV = st.s.V;
    st.I0_prev_used = st.s.I0;
    st.I1_prev_used = st.s.I1;
  }
  b.length = p.init_length - st.dl;
  b.width = p.init_width + st.dl;
  b.prev_volume = st.prevV;
  b.prevI[0] = st.I0_prev_used; b.prevI[1] = st.I1_prev_used; b.prevI[2] = st.I1_prev_used;
  b.com = st.s.com;
  b.prev_com = st.s.com;
  b.com_rate = st.s.com_rate;
  b.prev_com_rate = st.prev_com_rate;
  b.com_acc = st.com_acc;
}
SALP_HD void mixed_finish_dyn(const Motion32& s, Body64& b) {
  b.v[0] = s.v0; b.v[1] = s.v1; b.v[2] = s.v2;
  b.w[0] = s.w0; b.w[1] = s.w1; b.w[2] = s.w2;
  b.acc[0] = s.ac0; b.acc[1] = s.ac1; b.acc[2] = s.ac2;
  b.alp[0] = s.al0; b.alp[1] = s.al1; b.alp[2] = s.al2;
}
// max over the lanes of the warp that are still in the cycle loop.  The host build (tests/emu)
// returns "forever": it then runs the shape update after EVERY substep, the most adversarial
// neighbour a lane can have, so the CPU parity tests cover the no-op property relied on below.
SALP_HD int warp_max_int(int x) {
#ifdef __CUDA_ARCH__
  return __reduce_max_sync(__activemask(), x);
#elif defined(SALP_EMU_LANE_ONLY)
  return x;
#else
  (void)x;
  return 0x7fffffff;
#endif
}
// index of the next shape update after update j (0x7fffffff: none)
SALP_HD int next_update_after(int j, const PhasePlan& pp) {
  return (j < pp.upd_a_end || (j >= pp.upd_b_begin && j < pp.upd_b_end)) ? j + 1
         : (j < pp.upd_b_begin ? pp.upd_b_begin : 0x7fffffff);
}

template <bool NOISE, bool AXI>
SALP_HD int run_cycle_mixed(const SalpParams& p, const SalpDerived& dv, const CyclePlan& c, const double* time_table,
                            Body64& b, double& t_out, RandCtx* rc) {
  // ---- K: first k with !(t_k < total) in the dtype the reference compares in (robot.py:756) ----
  const int K = plan_substeps(c, time_table);
  t_out = 0.0;
  if (K <= 0) return K;            // K == 0: nothing moves; K < 0: range error (non-finite action)
  const PhasePlan pp = make_phase_plan(c, time_table, dv.inv_dt);
  const float dir[3] = {(float)c.dir[0], (float)c.dir[1], (float)c.dir[2]};

  ShapeTrack st;
  Coef32 g;
  Motion32 s;
  mixed_init_shape<AXI>(p, dv, b, dir, st, g);
  mixed_init_dyn(b, s);
  mixed_init_kin(b, s);

  // The kinematic update of substep k-1 only READS the (v, w) that the dynamics of substep k also
  // only reads, so the loop runs kin(k-1) side by side with dyn(k): two independent dependency
  // chains per iteration instead of one long one.  Substep 0's dynamics is peeled off in front,
  // substep K-1's kinematics behind.
  //
  // Shape updates: a lane needs them at j <= upd_a_end and upd_b_begin <= j <= upd_b_end.  Running
  // the update where the shape is static is a no-op (same shape -> same coefficients, all backward
  // differences 0), so the loop is split at the warp-uniform end W of the last lane's window:
  // loop A (j <= W) runs kin, dyn AND the update as one straight-line block -- three independent
  // chains for the scheduler, no divergent branch --, loop B (the coast, most substeps) runs the
  // lean body.  With the K-sort's second key (end of shape motion) W is close to every lane's own end.
  const int lane_end = pp.upd_a_end > pp.upd_b_end ? pp.upd_a_end : pp.upd_b_end;
  const int W = warp_max_int(lane_end < K ? lane_end : K);
  dyn_step<NOISE, false, AXI>(dv, g, s, rc, 0);
  shape_update<AXI>(p, dv, c, time_table, dir, 1, pp.k_T0, pp.k_jet, st, g);

  int k = 1;
  const int kA = W < K ? W : K;                  // part A covers updates j = k + 1 <= W
  double tj = time_table[1];                     // carried through part A by the same additions as the table
  // Chunk boundaries are FIXED (after iterations 16, 32, ...: kinematic updates 0..15, 16..31, ...)
  // so that the grouping of the fp32 chunk sums -- hence every bit of the result -- does not
  // depend on W, i.e. on which envs share the warp.
  while (k < K) {
    const int boundary = ((k - 1) & ~(SALP_MIXED_CHUNK - 1)) + SALP_MIXED_CHUNK + 1;
    const int cend = boundary < K ? boundary : K;
    const int aend = kA < cend ? kA : cend;
    for (; k < aend; k++) {
      kin_step(dv, s);
      dyn_step<NOISE, false, AXI>(dv, g, s, rc, k);
      tj = rn::dadd(tj, p.dt);
      shape_update_at<AXI>(p, dv, c, tj, dir, k + 1, pp.k_T0, pp.k_jet, st, g);
    }
    for (; k < cend; k++) {
      kin_step(dv, s);
      dyn_step<NOISE, true, AXI>(dv, g, s, rc, k);  // k >= W: the shape is static
    }
    if (k == boundary) flush_chunk(b, s);       // two-level sums (fp32 chunk partials -> fp64 totals)
  }
  // ---- the last substep's kinematic update ----
  kin_step(dv, s);
  flush_chunk(b, s);

  // ---- epilogue: back to the carried fp64 columns ----
  const double tK = time_table[K];
  mixed_finish_dyn(s, b);
  mixed_finish_shape(p, st, K, b);
  b.phase = phase_at(c, tK);
  b.speed_world = (double)sqrtf(s.vw0 * s.vw0 + s.vw1 * s.vw1);
  t_out = tK;
  return K;
}

template <>
SALP_HD int run_cycle<SALP_PRECISION_MIXED>(const SalpParams& p, const SalpDerived& dv, const CyclePlan& c,
                                            const double* time_table, Body64& b, double& t_out, RandCtx* rc) {
  (void)rc;
  // (warp-uniform: dv is a kernel argument; both forms give the same bits for axisymmetric parameters)
  if (dv.axisym) return run_cycle_mixed<false, true>(p, dv, c, time_table, b, t_out, nullptr);
  return run_cycle_mixed<false, false>(p, dv, c, time_table, b, t_out, nullptr);
}

// SalpParams.randomization != 0: per-env coefficient draws (Robot._randomize_parameters,
// robot.py:594-628: discharge coefficient, the two drag ratios and the added-mass diagonals, each
// mean * U(0.5, 1.5), re-drawn every cycle) replace the launch-wide constants in a thread-local copy
// of SalpDerived; OU disturbances run inside dyn_step.
SALP_HD void randomize_derived(const SalpParams& p, uint32_t flags, const RandCtx& rc, SalpDerived& k) {
  if (!(flags & SALP_RAND_DYNAMICS)) return;
  uint32_t r[16];
  for (uint32_t q = 0; q < 4; q++) rand_block(rc.seed, rc.gid, rc.episode, rc.cycle, SALP_RNG_DYNAMICS + q, r + 4 * q);
  const float cd = randomize_scalar((float)p.discharge_coefficient, 0.5f, r[0], 0.0f, 1.0f);
  k.ratio_f = randomize_scalar_default_bounds((float)p.drag_force_ratio, 0.5f, r[1]);
  k.torque_ratio = randomize_scalar_default_bounds((float)p.drag_torque_ratio, 0.5f, r[2]);
  k.jet_gain_f = (float)(-(double)cd * p.density / p.nozzle_area);
  k.jet_gain = (double)k.jet_gain_f;
  float ca[3], cat[3];
  for (int i = 0; i < 3; i++) {
    ca[i] = randomize_scalar((float)p.added_mass_force[i], 0.5f, r[3 + i], -1e30f, 1e30f);
    k.Car[i] = randomize_scalar((float)p.added_mass_rate_force[i], 0.5f, r[6 + i], -1e30f, 1e30f);
    cat[i] = randomize_scalar((float)p.added_mass_torque[i], 0.5f, r[9 + i], -1e30f, 1e30f);
  }
  for (int i = 0; i < 3; i++) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
    k.Ca[i] = ca[i];
    k.E[i] = 1.0f + ca[i];
    k.Cat[i] = cat[i];
    k.CaD[i] = ca[i2] - ca[i1];
    k.CatF[i] = 1.0f + cat[i];
  }
}

template <>
SALP_HD int run_cycle<SALP_PRECISION_MIXED_RANDOMIZED>(const SalpParams& p, const SalpDerived& dv, const CyclePlan& c,
                                                       const double* time_table, Body64& b, double& t_out,
                                                       RandCtx* rc) {
  SalpDerived k = dv;
  randomize_derived(p, (uint32_t)p.randomization, *rc, k);
  // (per-env coefficient draws are not axisymmetric: the general form)
  if (p.randomization & SALP_RAND_DISTURBANCE) return run_cycle_mixed<true, false>(p, k, c, time_table, b, t_out, rc);
  return run_cycle_mixed<false, false>(p, k, c, time_table, b, t_out, rc);
}
