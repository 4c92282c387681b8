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
//    yaw) are two-level sums: an fp32 partial per 16-substep chunk, flushed into an fp64 total.
//    sin/cos(yaw) = angle addition of the fp64-evaluated chunk base and the fp32 chunk partial.
//  * The substep count K and the phase of every substep are decided exactly as the reference
//    does (float32/float64 comparison quirks of SURVEY.md hard part 2) from the table t_k of
//    k-fold repeated `cycle_time += 0.01` additions.
#pragma once
#include "salp_loop_f64.cuh"

#define SALP_MIXED_CHUNK 16

// ---- MUFU-based reciprocal / norm with one Newton step (~1 ulp, no slow-path branches) --------
SALP_HD float fast_rcp(float x) {
#ifdef __CUDA_ARCH__
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return fmaf(r, fmaf(-x, r, 1.0f), r);
#else
  return 1.0f / x;
#endif
}
SALP_HD float fast_norm3(float a, float b, float c) {
  float s = fmaf(a, a, fmaf(b, b, c * c));
#ifdef __CUDA_ARCH__
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaxf(s, 1e-35f)));
  float y = s * r;
  float e = fmaf(-y, y, s);
  return fmaf(0.5f * r, e, y);
#else
  return sqrtf(s);
#endif
}

// np_sincosf without the separately-rounded steps: same Cody-Waite + minimax kernels (1 ulp),
// free to contract.  |x| <= 71476.  Used outside the substep loop and on the wide-angle path.
SALP_HD void sincos32(float x, float& sn, float& cs) {
  float q = (x * 0x1.45f306p-1f + 0x1.8p+23f) - 0x1.8p+23f;
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
  float dt, ratio_f, pi, end_aspect, inv_aspect_span, half_rho_neg, torque_ratio, arm0;
  float Ca[3], E[3], Cat[3], Car[3], CaD[3], CatF[3];
  float thi[3], tspan[3], rhi[3], rspan[3];
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
  k.dt = (float)p.dt;
  k.ratio_f = (float)p.drag_force_ratio;
  k.pi = (float)M_PI;
  const double init_aspect = p.init_length / p.init_width;
  const double end_aspect = (p.init_length - p.max_contraction) / (p.max_contraction + p.init_width);
  k.end_aspect = (float)end_aspect;
  k.inv_aspect_span = (float)(1.0 / (init_aspect - end_aspect));
  k.half_rho_neg = (float)(-0.5 * p.density);
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
  return k;
}

// fp32 coefficient set of one substep: everything the Newton/Euler equations need from the body
// shape, pre-divided by mass / inertia.  Loop-invariant while the shape is static.  With
//   M = m, Ca/Car/Cat the added-mass diagonals, E = 1 + Ca, J_i = I_i (1 + Cat_i):
//   a_i     = aj_i + v_i (kdm_i (|v| + ratio) - mrm_i) - Ca_i a_prev,i - (w x (E o v))_i + fict_i
//   alpha_i = tj_i + w_i (kqI_i |w| + klI_i) - Cat_i alpha_prev,i - w_i1 w_i2 JdI_i - v_i1 v_i2 AdI_i
// which is robot.py:789-851 / dynamics.py:6-174 with the common factors 1/m, 1/I_i cancelled
// (Coriolis and added-mass cross products merged; the two cross products of a vector with its own
// diagonal scaling collapse to one product per component).
struct Coef32 {
  float aj[3];        // F_jet / m                                       (robot.py:937-951)
  float kdm[3];       // -rho/2 area_i Ct_i / m                          (dynamics.py:111-116)
  float mrm[3];       // mass_rate Car_i / m                             (dynamics.py:139)
  float com, com_rate, com_acc;                                       // robot.py:898-922
  float tj1, tj2;     // (arm x F_jet)_i / I_i                           (robot.py:931-935)
  float kqI[3];       // -rho/2 Cr_i area_i dims_i / I_i                 (dynamics.py:120-128)
  float klI[3];       // (ratio -rho/2 Cr_i area_i width - I_rate_i) / I_i   (+ deform torque, :172-174)
  float JdI[3];       // (J_i2 - J_i1) / I_i
  float AdI[3];       // m (Ca_i2 - Ca_i1) / I_i
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

// All fp32 coefficients of the coming substep from the fp64 shape results.
SALP_HD void make_coefs(const SalpDerived& k, const float dir[3], bool jet_on, float lh, float wh, float m,
                        float I0, float I1, float I_rate0, float I_rate1, float mass_rate, float dV_dt,
                        float com, float com_rate, float com_acc, Coef32& g) {
  const float inv_m = fast_rcp(m);
  const float inv_I0 = fast_rcp(I0), inv_I1 = fast_rcp(I1);
  const float a0 = k.pi * wh * wh, a1 = k.pi * lh * wh;          // geometry.py:68-75
  float nr = (lh * fast_rcp(wh) - k.end_aspect) * k.inv_aspect_span;   // geometry.py:105-123
  nr = fminf(fmaxf(nr, 0.0f), 1.0f);
  const float width = wh + wh;
  const float w3 = width * width * width, l3 = 8.0f * lh * lh * lh;
  const float area[3] = {a0, a1, a1};
  const float dims[3] = {w3, l3, l3};
  const float I[3] = {I0, I1, I1};
  const float inv_I[3] = {inv_I0, inv_I1, inv_I1};
  const float I_rate[3] = {I_rate0, I_rate1, I_rate1};
  const float f = jet_on ? (float)k.jet_gain * dV_dt * dV_dt : 0.0f;
  const float armx = k.arm0 - lh;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
    float ct = k.thi[i] - nr * k.tspan[i];
    float cr = k.rhi[i] - nr * k.rspan[i];
    g.kdm[i] = k.half_rho_neg * area[i] * ct * inv_m;
    g.mrm[i] = mass_rate * k.Car[i] * inv_m;
    g.aj[i] = dir[i] * f * inv_m;
    float kr = k.half_rho_neg * cr * area[i];
    g.kqI[i] = kr * dims[i] * inv_I[i];
    g.klI[i] = (k.torque_ratio * kr * width - I_rate[i]) * inv_I[i];
    g.JdI[i] = (I[i2] * k.CatF[i2] - I[i1] * k.CatF[i1]) * inv_I[i];
    g.AdI[i] = m * k.CaD[i] * inv_I[i];
  }
  g.tj1 = -armx * (dir[2] * f) * inv_I1;
  g.tj2 = armx * (dir[1] * f) * inv_I1;
  g.com = com;
  g.com_rate = com_rate;
  g.com_acc = com_acc;
}

// first k in [0, SALP_MAX_SUBSTEPS] with !(t_k < x) (strict) or !(t_k <= x) (non-strict); the
// guess x/dt is within one or two entries of the answer, so this is a couple of table reads.
template <bool STRICT>
SALP_HD int first_k_past(const double* table, double x, double inv_dt) {
  if (!(x == x)) return 0;                                         // NaN: every comparison is False
  double g = x * inv_dt;
  int k = g < 0.0 ? 0 : (g > (double)SALP_MAX_SUBSTEPS ? SALP_MAX_SUBSTEPS : (int)g);
  while (k > 0 && !(STRICT ? table[k - 1] < x : table[k - 1] <= x)) k--;
  while (k < SALP_MAX_SUBSTEPS && (STRICT ? table[k] < x : table[k] <= x)) k++;
  return k;
}

// Body-to-world rotation by successive elementary rotations (dynamics.py:35-58: Rz Ry Rx v)
#define SALP_ROTATE_TO_WORLD()                                     \
  float u1 = cph * v1 - sph * v2, u2 = sph * v1 + cph * v2;        \
  float r0 = cth * v0 + sth * u2, vw2 = cth * u2 - sth * v0;       \
  vw0 = cps * r0 - sps * u1;                                       \
  vw1 = sps * r0 + cps * u1;

template <>
SALP_HD int run_cycle<SALP_PRECISION_MIXED>(const SalpParams& p, const SalpDerived& dv, const CyclePlan& c,
                                            const double* time_table, Body64& b, double& t_out) {
  const float dt = dv.dt;

  // ---- K: first k with !(t_k < total) in the dtype the reference compares in (robot.py:756) ----
  const int K = plan_substeps(c, time_table);
  t_out = 0.0;
  if (K <= 0) return K;            // K == 0: nothing moves; K < 0: range error (non-finite action)

  // ---- integer phase plan (robot.py:640-649 on the table t_j; update j follows substep j-1) ----
  //   phase_j = 0 for j < k_T0, 1 for k_T0 <= j < k_jet, 2/3 afterwards
  //   the shape moves at updates j <= k_ref (refill ramp and its end) and k_T0 <= j <= k_jet (jet
  //   and its end); two more updates flush the first/second backward differences
  const int k_ref = first_k_past<true>(time_table, c.refill, dv.inv_dt);
  const int k_T0 = first_k_past<false>(time_table, c.T0, dv.inv_dt);
  const int k_jet = first_k_past<false>(time_table, c.Tjet, dv.inv_dt);
  const int upd_a_end = (k_ref > 1 ? k_ref : 1) + 2;
  const int upd_b_begin = k_T0;
  const int upd_b_end = (k_jet > k_T0 ? k_jet : k_T0) + 2;
  const float dir[3] = {(float)c.dir[0], (float)c.dir[1], (float)c.dir[2]};

  // ---- prologue: shape-derived state of the first substep from the carried columns ----
  Shape64 s;
  Coef32 g;
  double prev_com_rate = b.prev_com_rate;
  double com_acc64 = b.com_acc;
  double prevV = b.prev_volume;
  double I0_prev_used, I1_prev_used;     // inertia used by the latest substep's Euler equations (robot.py:896)
  double dl = 0.0;
  {
    double lh = 0.5 * b.length, wh = 0.5 * b.width, wm, com_now;
    shape64_at(p, dv, lh, wh, s.V, s.I0, s.I1, com_now, wm);
    double dV_dt = (s.V - b.prev_volume) * dv.inv_dt;
    // the carried centre of mass may be stale w.r.t. length/width (Robot.reset quirk, robot.py:478)
    s.com = b.com;
    s.com_rate = b.com_rate;
    make_coefs(dv, dir, b.phase == 1, (float)lh, (float)wh, (float)(dv.m0 + wm), (float)s.I0, (float)s.I1,
               (float)((s.I0 - b.prevI[0]) * dv.inv_dt), (float)((s.I1 - b.prevI[1]) * dv.inv_dt),
               (float)(p.density * dV_dt), (float)dV_dt, (float)b.com, (float)b.com_rate, (float)b.com_acc, g);
    I0_prev_used = s.I0;
    I1_prev_used = s.I1;
  }
  int last_update = 0;

  // ---- fp32 motion state ----
  float v0 = (float)b.v[0], v1 = (float)b.v[1], v2 = (float)b.v[2];
  float w0 = (float)b.w[0], w1 = (float)b.w[1], w2 = (float)b.w[2];
  float ac0 = (float)b.acc[0], ac1 = (float)b.acc[1], ac2 = (float)b.acc[2];
  float al0 = (float)b.alp[0], al1 = (float)b.alp[1], al2 = (float)b.alp[2];
  float phi = (float)b.eul[0], theta = (float)b.eul[1];
  float sph, cph, sth, cth;
  sincos32(phi, sph, cph);
  sincos32(theta, sth, cth);
  double psi64 = b.eul[2];
  float sps, cps;
  {
    double sb64, cb64;
    sincos(psi64, &sb64, &cb64);
    sps = (float)sb64;
    cps = (float)cb64;
  }
  float psi_lo = 0.f, pw_lo0 = 0.f, pw_lo1 = 0.f, pw_lo2 = 0.f;
  float pos_lo0 = 0.f, pos_lo1 = 0.f, pos_lo2 = 0.f, ang_lo0 = 0.f, ang_lo1 = 0.f, ang_lo2 = 0.f;
  float vw0 = 0.f, vw1 = 0.f;
  // The kinematic update (Euler angles, world position, body-frame integrals; robot.py:864-875)
  // of substep k-1 only READS the (v, w) that the dynamics of substep k also only reads, so it
  // is issued one iteration late, side by side with the next substep's force/torque chains: two
  // independent dependency chains per iteration instead of one long one.  kdt = 0 turns the
  // (not yet due) kinematic update of iteration 0 into a no-op.
  float kdt = 0.0f;

  for (int k0 = 0; k0 < K; k0 += SALP_MIXED_CHUNK) {
    const int kend = k0 + SALP_MIXED_CHUNK < K ? k0 + SALP_MIXED_CHUNK : K;
    // roll / pitch move by < 0.1 rad per chunk; beyond 0.45 rad the chunk takes the range-reduced path
    const bool wide = fabsf(phi) > 0.45f || fabsf(theta) > 0.45f;
    for (int k = k0; k < kend; k++) {
      // ---- kinematics of the previous substep (robot.py:864-875) ----
      {
        float rcth = fast_rcp(cth);                              // dynamics.py:21-31 at the OLD roll/pitch
        float q = sph * w1 + cph * w2;
        float er0 = fmaf(sth * rcth, q, w0);
        float er1 = cph * w1 - sph * w2;
        float dpsi = (q * rcth) * kdt;
        phi = fmaf(er0, kdt, phi);
        theta = fmaf(er1, kdt, theta);
        psi_lo += dpsi;
        if (wide) { sincos32(phi, sph, cph); sincos32(theta, sth, cth); }
        else { sincos_small(phi, sph, cph); sincos_small(theta, sth, cth); }
        float sd_, cd_;
        sincos_small(dpsi, sd_, cd_);                            // yaw: rotate (sin, cos) by the increment
        float ns = sps * cd_ + cps * sd_;
        cps = cps * cd_ - sps * sd_;
        sps = ns;
        SALP_ROTATE_TO_WORLD();
        pw_lo0 = fmaf(vw0, kdt, pw_lo0); pw_lo1 = fmaf(vw1, kdt, pw_lo1); pw_lo2 = fmaf(vw2, kdt, pw_lo2);
        pos_lo0 = fmaf(v0, kdt, pos_lo0); pos_lo1 = fmaf(v1, kdt, pos_lo1); pos_lo2 = fmaf(v2, kdt, pos_lo2);
        ang_lo0 = fmaf(w0, kdt, ang_lo0); ang_lo1 = fmaf(w1, kdt, ang_lo1); ang_lo2 = fmaf(w2, kdt, ang_lo2);
        kdt = dt;
      }
      // ---- _newton_equations (robot.py:789-823) ----
      float sd = fast_norm3(v0, v1, v2) + dv.ratio_f;            // |v| v + ratio v = v (|v| + ratio)
      float ev0 = dv.E[0] * v0, ev1 = dv.E[1] * v1, ev2 = dv.E[2] * v2;
      float t1 = w2 * g.com, t2 = -w1 * g.com;                   // w x c, c = (com, 0, 0)   robot.py:806-810
      float fict0 = (w1 * t2 - w2 * t1) + g.com_acc;
      float fict1 = fmaf(al2, g.com, fmaf(2.0f * w2, g.com_rate, -w0 * t2));
      float fict2 = fmaf(-al1, g.com, fmaf(-2.0f * w1, g.com_rate, w0 * t1));
      float na0 = g.aj[0] + v0 * fmaf(g.kdm[0], sd, -g.mrm[0]) - dv.Ca[0] * ac0 - (w1 * ev2 - w2 * ev1) + fict0;
      float na1 = g.aj[1] + v1 * fmaf(g.kdm[1], sd, -g.mrm[1]) - dv.Ca[1] * ac1 - (w2 * ev0 - w0 * ev2) + fict1;
      float na2 = g.aj[2] + v2 * fmaf(g.kdm[2], sd, -g.mrm[2]) - dv.Ca[2] * ac2 - (w0 * ev1 - w1 * ev0) + fict2;
      // ---- _euler_equations (robot.py:825-851) ----
      float wn = fast_norm3(w0, w1, w2);
      float nl0 = w0 * fmaf(g.kqI[0], wn, g.klI[0]) - dv.Cat[0] * al0 - (w1 * w2) * g.JdI[0] - (v1 * v2) * g.AdI[0];
      float nl1 = g.tj1 + w1 * fmaf(g.kqI[1], wn, g.klI[1]) - dv.Cat[1] * al1 - (w2 * w0) * g.JdI[1] - (v2 * v0) * g.AdI[1];
      float nl2 = g.tj2 + w2 * fmaf(g.kqI[2], wn, g.klI[2]) - dv.Cat[2] * al2 - (w0 * w1) * g.JdI[2] - (v0 * v1) * g.AdI[2];
      ac0 = na0; ac1 = na1; ac2 = na2;
      al0 = nl0; al1 = nl1; al2 = nl2;
      // ---- _update_motion_states, velocities (robot.py:861-862) ----
      v0 = fmaf(ac0, dt, v0); v1 = fmaf(ac1, dt, v1); v2 = fmaf(ac2, dt, v2);
      w0 = fmaf(al0, dt, w0); w1 = fmaf(al1, dt, w1); w2 = fmaf(al2, dt, w2);

      // ---- cycle_time += dt; update_state; update_properties (robot.py:674-678, 640-668) ----
      const int j = k + 1;
      if (j <= upd_a_end || (j >= upd_b_begin && j <= upd_b_end)) {
        const double t = time_table[j];
        const int phase = j < k_T0 ? 0 : (j < k_jet ? 1 : 2);
        dl = shape_delta(phase, t, c.refill, c.T0, (double)c.contraction32, c.contract_rate, c.release_rate);
        double lh = 0.5 * (p.init_length - dl), wh = 0.5 * (p.init_width + dl);
        double V, I0n, I1n, com, wm;
        shape64_at(p, dv, lh, wh, V, I0n, I1n, com, wm);
        double dV_dt = (V - s.V) * dv.inv_dt;
        double com_rate = (com - s.com) * dv.inv_dt;                // robot.py:901-910
        com_acc64 = (com_rate - prev_com_rate) * dv.inv_dt;         // robot.py:912-922
        prev_com_rate = com_rate;
        make_coefs(dv, dir, phase == 1, (float)lh, (float)wh, (float)(dv.m0 + wm), (float)I0n, (float)I1n,
                   (float)((I0n - s.I0) * dv.inv_dt), (float)((I1n - s.I1) * dv.inv_dt),
                   (float)(p.density * dV_dt), (float)dV_dt, (float)com, (float)com_rate, (float)com_acc64, g);
        prevV = s.V;
        I0_prev_used = s.I0;
        I1_prev_used = s.I1;
        s.V = V; s.I0 = I0n; s.I1 = I1n; s.com = com; s.com_rate = com_rate;
        last_update = j;
      }
    }
    // two-level sums: fold the fp32 chunk partials into the fp64 totals, re-anchor sin/cos(yaw)
    b.pw[0] += (double)pw_lo0; b.pw[1] += (double)pw_lo1; b.pw[2] += (double)pw_lo2;
    b.pos[0] += (double)pos_lo0; b.pos[1] += (double)pos_lo1; b.pos[2] += (double)pos_lo2;
    b.ang[0] += (double)ang_lo0; b.ang[1] += (double)ang_lo1; b.ang[2] += (double)ang_lo2;
    psi64 += (double)psi_lo;
    {
      double sb64, cb64;
      sincos(psi64, &sb64, &cb64);
      sps = (float)sb64;
      cps = (float)cb64;
    }
    psi_lo = 0.f; pw_lo0 = pw_lo1 = pw_lo2 = 0.f;
    pos_lo0 = pos_lo1 = pos_lo2 = 0.f; ang_lo0 = ang_lo1 = ang_lo2 = 0.f;
  }
  // ---- the last substep's kinematic update (pipelined one iteration late) ----
  {
    float rcth = fast_rcp(cth);
    float q = sph * w1 + cph * w2;
    float er0 = fmaf(sth * rcth, q, w0);
    float er1 = cph * w1 - sph * w2;
    float dpsi = (q * rcth) * dt;
    phi = fmaf(er0, dt, phi);
    theta = fmaf(er1, dt, theta);
    sincos32(phi, sph, cph);
    sincos32(theta, sth, cth);
    psi64 += (double)dpsi;
    double sb64, cb64;
    sincos(psi64, &sb64, &cb64);
    sps = (float)sb64;
    cps = (float)cb64;
    SALP_ROTATE_TO_WORLD();
    b.pw[0] += (double)(vw0 * dt); b.pw[1] += (double)(vw1 * dt); b.pw[2] += (double)(vw2 * dt);
    b.pos[0] += (double)(v0 * dt); b.pos[1] += (double)(v1 * dt); b.pos[2] += (double)(v2 * dt);
    b.ang[0] += (double)(w0 * dt); b.ang[1] += (double)(w1 * dt); b.ang[2] += (double)(w2 * dt);
  }

  // ---- epilogue: back to the carried fp64 columns ----
  if (last_update != K) {        // static tail: update_properties re-assigned the same shape (robot.py:651-668)
    prevV = s.V;
    I0_prev_used = s.I0;
    I1_prev_used = s.I1;
  }
  const double tK = time_table[K];
  b.v[0] = v0; b.v[1] = v1; b.v[2] = v2;
  b.w[0] = w0; b.w[1] = w1; b.w[2] = w2;
  b.acc[0] = ac0; b.acc[1] = ac1; b.acc[2] = ac2;
  b.alp[0] = al0; b.alp[1] = al1; b.alp[2] = al2;
  b.eul[0] = phi; b.eul[1] = theta; b.eul[2] = psi64;
  b.phase = phase_at(c, tK);
  b.length = p.init_length - dl;
  b.width = p.init_width + dl;
  b.prev_volume = prevV;
  b.prevI[0] = I0_prev_used; b.prevI[1] = I1_prev_used; b.prevI[2] = I1_prev_used;
  b.com = s.com;
  b.prev_com = s.com;
  b.com_rate = s.com_rate;
  b.prev_com_rate = prev_com_rate;
  b.com_acc = com_acc64;
  b.speed_world = (double)sqrtf(vw0 * vw0 + vw1 * vw1);
  t_out = tK;
  return K;
}
