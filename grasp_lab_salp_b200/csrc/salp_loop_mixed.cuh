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

// np_sincosf without the separately-rounded steps: same Cody-Waite + minimax kernels (1 ulp),
// free to contract.  |x| <= 71476.
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

// fp32 coefficient set of one substep: everything the Newton/Euler equations need from the body
// shape.  Loop-invariant while the shape is static.
struct Coef32 {
  float m, inv_m;
  float kd[3];        // -rho/2 * area_i * Ct_i                      (drag force, dynamics.py:111-116)
  float kq[3];        // -rho/2 * Cr_i * area_i * dims_i             (quadratic drag torque, :120-128)
  float kl[3];        // ratio * -rho/2 * Cr_i * area_i * width      (linear drag torque)
  float I[2], inv_I[2];   // I[1] == I[2] (geometry.py:134-183)
  float I_rate[2];    // (I - prev_I)/dt                              (robot.py:888-896)
  float mass_rate;    // (m_w - m_w,prev)/dt                          (geometry.py:98-101)
  float com, com_rate, com_acc;                                    // robot.py:898-922
  float Fj[3];        // jet force (depends on the shape sequence only; robot.py:937-951)
  float Tj1, Tj2;     // jet torque arm x Fj                          (robot.py:931-935)
};

// fp64 side of the shape: the quantities that are differenced.
struct Shape64 {
  double V;           // water volume (ellipsoid - tube), robot.py:1055-1056
  double I0, I1;      // inertia diagonal
  double com;         // centre of mass x
  double com_rate;
};

struct ShapeConst {   // per-launch constants of the fp64 chain (from SalpParams)
  double inv_dt, four_thirds_pi, skin3, c2, c1, c0, comA, comB, mtot0, m0;
};

SALP_HD ShapeConst make_shape_const(const SalpParams& p) {
  ShapeConst k;
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
  return k;
}

// fp64 shape chain at half-length lh, half-width wh: 19 flop + 1 division.
SALP_HD void shape64_at(const SalpParams& p, const ShapeConst& k, double lh, double wh, double& V,
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

// fp32 coefficients that depend on the shape but are never differenced.
SALP_HD void shape32_coefs(const SalpParams& p, float lh, float wh, float m, float I0, float I1, Coef32& g) {
  const float pi = (float)M_PI;
  float a0 = pi * wh * wh, a1 = pi * lh * wh;            // geometry.py:68-75
  // geometry.py:105-123
  float aspect = lh / wh;
  const float init_aspect = (float)(p.init_length / p.init_width);
  const float end_aspect = (float)((p.init_length - p.max_contraction) / (p.max_contraction + p.init_width));
  float nr = (aspect - end_aspect) * (1.0f / (init_aspect - end_aspect));
  nr = fminf(fmaxf(nr, 0.0f), 1.0f);
  const float hr = (float)(-0.5 * p.density);
  float width = wh + wh;
  float w3 = width * width * width, l3 = 8.0f * lh * lh * lh;
  const float area[3] = {a0, a1, a1};
  const float dims[3] = {w3, l3, l3};
#pragma unroll
  for (int i = 0; i < 3; i++) {
    float thi = (float)p.trans_drag_range[2 * i + 1], tlo = (float)p.trans_drag_range[2 * i];
    float rhi = (float)p.rot_drag_range[2 * i + 1], rlo = (float)p.rot_drag_range[2 * i];
    float ct = thi - nr * (thi - tlo);
    float cr = rhi - nr * (rhi - rlo);
    g.kd[i] = hr * area[i] * ct;
    float kr = hr * cr * area[i];
    g.kq[i] = kr * dims[i];
    g.kl[i] = (float)p.drag_torque_ratio * kr * width;
  }
  g.m = m;
  g.inv_m = 1.0f / m;
  g.I[0] = I0;
  g.I[1] = I1;
  g.inv_I[0] = 1.0f / I0;
  g.inv_I[1] = 1.0f / I1;
}

// jet force / torque of the coming substep (robot.py:931-951, dynamics.py:88-107)
SALP_HD void jet32(const SalpParams& p, const CyclePlan& c, int phase, double dV_dt, double mass_rate,
                   float lh, Coef32& g) {
  float f = 0.0f;
  if (phase == 1) f = (float)(-p.discharge_coefficient * (mass_rate * (dV_dt / p.nozzle_area)));
  g.Fj[0] = (float)c.dir[0] * f;
  g.Fj[1] = (float)c.dir[1] * f;
  g.Fj[2] = (float)c.dir[2] * f;
  float armx = -(float)(p.nozzle_length1 + p.nozzle_length2) - lh;
  g.Tj1 = -armx * g.Fj[2];
  g.Tj2 = armx * g.Fj[1];
}

template <>
SALP_HD int run_cycle<SALP_PRECISION_MIXED>(const SalpParams& p, const CyclePlan& c, const double* time_table,
                                            Body64& b, double& t_out) {
  const ShapeConst sc = make_shape_const(p);
  const float dt = (float)p.dt;

  // ---- K: first k with !(t_k < total) in the dtype the reference compares in (robot.py:756) ----
  const int K = plan_substeps(c, time_table);
  t_out = 0.0;
  if (K <= 0) return K;            // K == 0: nothing moves; K < 0: range error (non-finite action)

  // ---- prologue: shape-derived state of the first substep from the carried columns ----
  Shape64 s;
  Coef32 g;
  double I0_prev_used, I1_prev_used;     // inertia used by the latest substep's Euler equations (robot.py:896)
  double prev_com_rate = b.prev_com_rate;
  double dl = 0.0;
  bool first = true;                      // forces a shape update after the first substep
  int settle = 0;
  int phase = b.phase;
  {
    double lh = 0.5 * b.length, wh = 0.5 * b.width, wm, com_now;
    shape64_at(p, sc, lh, wh, s.V, s.I0, s.I1, com_now, wm);
    double dV_dt = (s.V - b.prev_volume) * sc.inv_dt;
    double mass_rate = p.density * dV_dt;
    shape32_coefs(p, (float)lh, (float)wh, (float)(sc.m0 + wm), (float)s.I0, (float)s.I1, g);
    g.mass_rate = (float)mass_rate;
    g.I_rate[0] = (float)((s.I0 - b.prevI[0]) * sc.inv_dt);
    g.I_rate[1] = (float)((s.I1 - b.prevI[1]) * sc.inv_dt);
    // the carried centre of mass may be stale w.r.t. length/width (Robot.reset quirk, robot.py:478)
    s.com = b.com;
    s.com_rate = b.com_rate;
    g.com = (float)b.com;
    g.com_rate = (float)b.com_rate;
    g.com_acc = (float)b.com_acc;
    jet32(p, c, phase, dV_dt, mass_rate, (float)lh, g);
    I0_prev_used = s.I0;
    I1_prev_used = s.I1;
  }
  double com_acc64 = b.com_acc;
  double prevV = b.prev_volume;

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
  double sb64, cb64;
  sincos(psi64, &sb64, &cb64);
  float sb = (float)sb64, cb = (float)cb64;
  float psi_lo = 0.f, pw_lo0 = 0.f, pw_lo1 = 0.f, pw_lo2 = 0.f;
  float pos_lo0 = 0.f, pos_lo1 = 0.f, pos_lo2 = 0.f, ang_lo0 = 0.f, ang_lo1 = 0.f, ang_lo2 = 0.f;
  float vw0 = 0.f, vw1 = 0.f;
  const float Ca0 = (float)p.added_mass_force[0], Ca1 = (float)p.added_mass_force[1], Ca2 = (float)p.added_mass_force[2];
  const float Car0 = (float)p.added_mass_rate_force[0], Car1 = (float)p.added_mass_rate_force[1],
              Car2 = (float)p.added_mass_rate_force[2];
  const float Cat0 = (float)p.added_mass_torque[0], Cat1 = (float)p.added_mass_torque[1],
              Cat2 = (float)p.added_mass_torque[2];
  const float ratio_f = (float)p.drag_force_ratio;

  for (int k = 0; k < K; k++) {
    // ---- _newton_equations (robot.py:789-823) ----
    const float m = g.m;
    float vn = sqrtf(v0 * v0 + v1 * v1 + v2 * v2);
    float sd = vn + ratio_f;                                   // |v| v + ratio v = v (|v| + ratio)
    float Fd0 = g.kd[0] * v0 * sd, Fd1 = g.kd[1] * v1 * sd, Fd2 = g.kd[2] * v2 * sd;
    float mv0 = m * v0, mv1 = m * v1, mv2 = m * v2;
    float Fc0 = w2 * mv1 - w1 * mv2, Fc1 = w0 * mv2 - w2 * mv0, Fc2 = w1 * mv0 - w0 * mv1;   // -w x (M v)
    float am0 = m * Ca0, am1 = m * Ca1, am2 = m * Ca2;
    float av0 = am0 * v0, av1 = am1 * v1, av2 = am2 * v2;
    float Fa0 = -(am0 * ac0 + (w1 * av2 - w2 * av1) + (g.mass_rate * Car0) * v0);
    float Fa1 = -(am1 * ac1 + (w2 * av0 - w0 * av2) + (g.mass_rate * Car1) * v1);
    float Fa2 = -(am2 * ac2 + (w0 * av1 - w1 * av0) + (g.mass_rate * Car2) * v2);
    // fictitious forces of the moving centre of mass c = (com, 0, 0)   robot.py:806-810
    float t1 = w2 * g.com, t2 = -w1 * g.com;
    float cen0 = w1 * t2 - w2 * t1, cen1 = -w0 * t2, cen2 = w0 * t1;
    float cor1 = 2.0f * (w2 * g.com_rate), cor2 = -2.0f * (w1 * g.com_rate);
    float tan1 = al2 * g.com, tan2 = -al1 * g.com;
    float Ff0 = m * (cen0 + g.com_acc), Ff1 = m * (cen1 + cor1 + tan1), Ff2 = m * (cen2 + cor2 + tan2);
    float na0 = (g.Fj[0] + Fd0 + Fa0 + Fc0 + Ff0) * g.inv_m;
    float na1 = (g.Fj[1] + Fd1 + Fa1 + Fc1 + Ff1) * g.inv_m;
    float na2 = (g.Fj[2] + Fd2 + Fa2 + Fc2 + Ff2) * g.inv_m;

    // ---- _euler_equations (robot.py:825-851) ----
    const float I0 = g.I[0], I1 = g.I[1];
    float Iw0 = I0 * w0, Iw1 = I1 * w1, Iw2 = I1 * w2;
    float Tc0 = w2 * Iw1 - w1 * Iw2, Tc1 = w0 * Iw2 - w2 * Iw0, Tc2 = w1 * Iw0 - w0 * Iw1;   // -w x (I w)
    float wn = sqrtf(w0 * w0 + w1 * w1 + w2 * w2);
    float Td0 = w0 * (g.kq[0] * wn + g.kl[0]), Td1 = w1 * (g.kq[1] * wn + g.kl[1]), Td2 = w2 * (g.kq[2] * wn + g.kl[2]);
    float Tdf0 = -(g.I_rate[0] * w0), Tdf1 = -(g.I_rate[1] * w1), Tdf2 = -(g.I_rate[1] * w2);
    float at0 = I0 * Cat0, at1 = I1 * Cat1, at2 = I1 * Cat2;
    float aw0 = at0 * w0, aw1 = at1 * w1, aw2 = at2 * w2;
    float Ta0 = -(at0 * al0 + (w1 * aw2 - w2 * aw1) + (v1 * av2 - v2 * av1));
    float Ta1 = -(at1 * al1 + (w2 * aw0 - w0 * aw2) + (v2 * av0 - v0 * av2));
    float Ta2 = -(at2 * al2 + (w0 * aw1 - w1 * aw0) + (v0 * av1 - v1 * av0));
    float nl0 = (Td0 + Tc0 + Tdf0 + Ta0) * g.inv_I[0];
    float nl1 = (g.Tj1 + Td1 + Tc1 + Tdf1 + Ta1) * g.inv_I[1];
    float nl2 = (g.Tj2 + Td2 + Tc2 + Tdf2 + Ta2) * g.inv_I[1];
    ac0 = na0; ac1 = na1; ac2 = na2;
    al0 = nl0; al1 = nl1; al2 = nl2;

    // ---- _update_motion_states (robot.py:860-875) ----
    v0 += ac0 * dt; v1 += ac1 * dt; v2 += ac2 * dt;
    w0 += al0 * dt; w1 += al1 * dt; w2 += al2 * dt;
    float rcth = 1.0f / cth;                                   // dynamics.py:21-31 at the OLD roll/pitch
    float tth = sth * rcth;
    float q = sph * w1 + cph * w2;
    float er0 = w0 + tth * q;
    float er1 = cph * w1 - sph * w2;
    float er2 = q * rcth;
    phi += er0 * dt;
    theta += er1 * dt;
    psi_lo += er2 * dt;
    sincos32(phi, sph, cph);
    sincos32(theta, sth, cth);
    float sl, cl;
    sincos32(psi_lo, sl, cl);
    float sps = sb * cl + cb * sl, cps = cb * cl - sb * sl;
    // dynamics.py:35-58: R = Rz(psi) Ry(theta) Rx(phi)
    float r00 = cps * cth, r01 = cps * sth * sph - sps * cph, r02 = cps * sth * cph + sps * sph;
    float r10 = sps * cth, r11 = sps * sth * sph + cps * cph, r12 = sps * sth * cph - cps * sph;
    float r20 = -sth, r21 = cth * sph, r22 = cth * cph;
    vw0 = r00 * v0 + r01 * v1 + r02 * v2;
    vw1 = r10 * v0 + r11 * v1 + r12 * v2;
    float vw2 = r20 * v0 + r21 * v1 + r22 * v2;
    pw_lo0 += vw0 * dt; pw_lo1 += vw1 * dt; pw_lo2 += vw2 * dt;
    pos_lo0 += v0 * dt; pos_lo1 += v1 * dt; pos_lo2 += v2 * dt;
    ang_lo0 += w0 * dt; ang_lo1 += w1 * dt; ang_lo2 += w2 * dt;

    // two-level sums: fold the fp32 chunk partials into the fp64 totals
    if ((k & (SALP_MIXED_CHUNK - 1)) == SALP_MIXED_CHUNK - 1 || k == K - 1) {
      b.pw[0] += (double)pw_lo0; b.pw[1] += (double)pw_lo1; b.pw[2] += (double)pw_lo2;
      b.pos[0] += (double)pos_lo0; b.pos[1] += (double)pos_lo1; b.pos[2] += (double)pos_lo2;
      b.ang[0] += (double)ang_lo0; b.ang[1] += (double)ang_lo1; b.ang[2] += (double)ang_lo2;
      psi64 += (double)psi_lo;
      sincos(psi64, &sb64, &cb64);
      sb = (float)sb64; cb = (float)cb64;
      psi_lo = 0.f; pw_lo0 = pw_lo1 = pw_lo2 = 0.f;
      pos_lo0 = pos_lo1 = pos_lo2 = 0.f; ang_lo0 = ang_lo1 = ang_lo2 = 0.f;
    }

    // ---- cycle_time += dt; update_state; update_properties (robot.py:674-678, 640-668) ----
    const double t = time_table[k + 1];
    phase = phase_at(c, t);
    const double dl_new = shape_delta(phase, t, c.refill, c.T0, (double)c.contraction32, c.contract_rate,
                                      c.release_rate);
    const bool changed = first || dl_new != dl;
    first = false;
    if (changed || settle > 0) {
      settle = changed ? 2 : settle - 1;
      dl = dl_new;
      double lh = 0.5 * (p.init_length - dl), wh = 0.5 * (p.init_width + dl);
      double V, I0n, I1n, com, wm;
      shape64_at(p, sc, lh, wh, V, I0n, I1n, com, wm);
      prevV = s.V;
      double dV_dt = (V - s.V) * sc.inv_dt;
      double mass_rate = p.density * dV_dt;
      I0_prev_used = s.I0;
      I1_prev_used = s.I1;
      double com_rate = (com - s.com) * sc.inv_dt;                // robot.py:901-910
      com_acc64 = (com_rate - prev_com_rate) * sc.inv_dt;         // robot.py:912-922
      prev_com_rate = com_rate;
      shape32_coefs(p, (float)lh, (float)wh, (float)(sc.m0 + wm), (float)I0n, (float)I1n, g);
      g.mass_rate = (float)mass_rate;
      g.I_rate[0] = (float)((I0n - s.I0) * sc.inv_dt);
      g.I_rate[1] = (float)((I1n - s.I1) * sc.inv_dt);
      g.com = (float)com;
      g.com_rate = (float)com_rate;
      g.com_acc = (float)com_acc64;
      jet32(p, c, phase, dV_dt, mass_rate, (float)lh, g);
      s.V = V; s.I0 = I0n; s.I1 = I1n; s.com = com; s.com_rate = com_rate;
    } else {
      // static shape, all differences already flushed: only the jet switch can change
      // (JET -> COAST with an unchanged shape cannot happen: the jet always moves the shape)
      prevV = s.V;
      I0_prev_used = s.I0;
      I1_prev_used = s.I1;
    }
  }

  // ---- epilogue: back to the carried fp64 columns ----
  b.v[0] = v0; b.v[1] = v1; b.v[2] = v2;
  b.w[0] = w0; b.w[1] = w1; b.w[2] = w2;
  b.acc[0] = ac0; b.acc[1] = ac1; b.acc[2] = ac2;
  b.alp[0] = al0; b.alp[1] = al1; b.alp[2] = al2;
  b.eul[0] = phi; b.eul[1] = theta; b.eul[2] = psi64;
  b.phase = phase;
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
  t_out = time_table[K];
  return K;
}
