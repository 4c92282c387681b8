// salp_loop_f64.cuh -- SALP_PRECISION_F64: the reference-mode substep loop.
//
// A quirk-for-quirk float64 restatement of Robot.step() (robot.py:670-678) for one env held in
// registers: same order of updates, same lag-by-one quantities (a_prev, alpha_prev, the phase
// decided at the end of the previous substep), same side-effecting getters (prev_I is refreshed
// by the deform-torque call, so the added-mass torque sees I_rate == 0), same finite
// differences.  Diagonal 3x3 matrices are 3-vectors; the centre of mass has only an x component.
// This loop is the in-repo "fast oracle" for the mixed-precision loop.
#pragma once
#include "salp_device.cuh"

struct Body64 {
  // motion state (robot.py:358-374)
  double v[3], w[3], eul[3], pw[3], acc[3], alp[3], pos[3], ang[3];
  double vwx, vwy;            // velocity_world[0:2] of the last substep run in this cycle
  double speed_world;         // |velocity_world[0:2]| carried across cycles (a K = 0 cycle keeps it)
  // geometry tail (robot.py:325-339)
  int phase;
  double length, width, volume, prev_volume, water_mass, mass, mass_rate;
  double area[3], I[3], prevI[3], ct[3], cr[3];
  double com, prev_com, com_rate, prev_com_rate, com_acc;
};

// everything update_properties derives from (length, width) alone (robot.py:658-660, 667-668)
SALP_DEV void refresh_shape_f64(const SalpParams& p, Body64& b) {
  cross_sections(b.length, b.width, b.area);
  b.volume = ellipsoid_volume(b.length, b.width) - p.tube_volume;
  b.water_mass = p.density * b.volume;
  b.mass = p.dry_mass + b.water_mass + p.nozzle_mass;
  double nr = drag_interp_weight(p, b.length, b.width);
#pragma unroll
  for (int i = 0; i < 3; i++) {
    b.ct[i] = p.trans_drag_range[2 * i + 1] - nr * (p.trans_drag_range[2 * i + 1] - p.trans_drag_range[2 * i]);
    b.cr[i] = p.rot_drag_range[2 * i + 1] - nr * (p.rot_drag_range[2 * i + 1] - p.rot_drag_range[2 * i]);
  }
  inertia_diag(b.length, b.width, p.nozzle_mass, b.I);
}

// One Robot.step(): update_dynamics, cycle_time += dt, update_state, update_properties.
SALP_DEV void substep_f64(const SalpParams& p, const CyclePlan& c, Body64& b, double& t) {
  const double dt = p.dt;
  const double half_rho = -0.5 * p.density;
  double* v = b.v;
  double* w = b.w;
  const double m = b.mass;
  // ---- _newton_equations (robot.py:789-823) ----
  // coriolis force -w x (M v)                                  dynamics.py:160-162
  double mv0 = m * v[0], mv1 = m * v[1], mv2 = m * v[2];
  double Fc0 = -(w[1] * mv2 - w[2] * mv1), Fc1 = -(w[2] * mv0 - w[0] * mv2), Fc2 = -(w[0] * mv1 - w[1] * mv0);
  // drag force                                                 dynamics.py:111-116
  double vn = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  double Fd[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    double k = half_rho * b.area[i] * b.ct[i];
    Fd[i] = k * vn * v[i] + p.drag_force_ratio * (k * v[i]);
  }
  // jet force, only in the JET phase decided at the end of the previous substep
  //                                                            robot.py:937-951, dynamics.py:88-101
  double Fj[3] = {0.0, 0.0, 0.0};
  if (b.phase == 1) {
    double jet_speed = ((b.volume - b.prev_volume) / dt) / p.nozzle_area;
#pragma unroll
    for (int i = 0; i < 3; i++) Fj[i] = -p.discharge_coefficient * (b.mass_rate * (c.dir[i] * jet_speed));
  }
  // added mass force, uses the previous substep's acceleration dynamics.py:132-141
  double am0 = m * p.added_mass_force[0], am1 = m * p.added_mass_force[1], am2 = m * p.added_mass_force[2];
  double av0 = am0 * v[0], av1 = am1 * v[1], av2 = am2 * v[2];
  double Fa0 = -(am0 * b.acc[0] + (w[1] * av2 - w[2] * av1) + (b.mass_rate * p.added_mass_rate_force[0]) * v[0]);
  double Fa1 = -(am1 * b.acc[1] + (w[2] * av0 - w[0] * av2) + (b.mass_rate * p.added_mass_rate_force[1]) * v[1]);
  double Fa2 = -(am2 * b.acc[2] + (w[0] * av1 - w[1] * av0) + (b.mass_rate * p.added_mass_rate_force[2]) * v[2]);
  // fictitious forces of the moving centre of mass c = (com, 0, 0)  robot.py:806-810
  double t1 = w[2] * b.com, t2 = -w[1] * b.com;               // w x c = (0, t1, t2)
  double cen0 = w[1] * t2 - w[2] * t1, cen1 = -w[0] * t2, cen2 = w[0] * t1;
  double cor1 = 2.0 * (w[2] * b.com_rate), cor2 = 2.0 * (-w[1] * b.com_rate);
  double tan1 = b.alp[2] * b.com, tan2 = -b.alp[1] * b.com;   // alpha_prev x c
  double Ff0 = m * (cen0 + b.com_acc), Ff1 = m * (cen1 + cor1 + tan1), Ff2 = m * (cen2 + cor2 + tan2);
  double a0 = (Fj[0] + Fd[0] + Fa0 + Fc0 + Ff0) / m;         // dynamics.py:6-10
  double a1 = (Fj[1] + Fd[1] + Fa1 + Fc1 + Ff1) / m;
  double a2 = (Fj[2] + Fd[2] + Fa2 + Fc2 + Ff2) / m;

  // ---- _euler_equations (robot.py:825-851) ----
  const double* I = b.I;
  double Iw0 = I[0] * w[0], Iw1 = I[1] * w[1], Iw2 = I[2] * w[2];
  double Tc0 = -(w[1] * Iw2 - w[2] * Iw1), Tc1 = -(w[2] * Iw0 - w[0] * Iw2), Tc2 = -(w[0] * Iw1 - w[1] * Iw0);
  double wn = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);   // dynamics.py:120-128
  double w3 = b.width * b.width * b.width, l3 = b.length * b.length * b.length;
  double dims[3] = {w3, l3, l3};
  double Td[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    double k = half_rho * b.cr[i] * b.area[i];
    Td[i] = k * wn * w[i] * dims[i] + p.drag_torque_ratio * (k * w[i] * b.width);
  }
  // jet torque arm x F, arm = (-(l1+l2) - length/2, 0, 0)      robot.py:931-935, geometry.py:127-130
  double armx = -(p.nozzle_length1 + p.nozzle_length2) + (-b.length / 2.0);
  double Tj1 = -armx * Fj[2], Tj2 = armx * Fj[1];
  // deform torque -I_rate w; get_inertia_matrix_rate() refreshes prev_I (robot.py:888-896), so the
  // added-mass torque below sees I_rate == 0
  double Tdf[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    Tdf[i] = -(((I[i] - b.prevI[i]) / dt) * w[i]);
    b.prevI[i] = I[i];
  }
  // added mass torque                                          dynamics.py:145-156
  double at0 = I[0] * p.added_mass_torque[0], at1 = I[1] * p.added_mass_torque[1], at2 = I[2] * p.added_mass_torque[2];
  double aw0 = at0 * w[0], aw1 = at1 * w[1], aw2 = at2 * w[2];
  double Ta0 = -(at0 * b.alp[0] + (w[1] * aw2 - w[2] * aw1) + (v[1] * av2 - v[2] * av1));
  double Ta1 = -(at1 * b.alp[1] + (w[2] * aw0 - w[0] * aw2) + (v[2] * av0 - v[0] * av2));
  double Ta2 = -(at2 * b.alp[2] + (w[0] * aw1 - w[1] * aw0) + (v[0] * av1 - v[1] * av0));
  double al0 = (0.0 + Td[0] + Tc0 + Tdf[0] + Ta0) / I[0];    // dynamics.py:14-17
  double al1 = (Tj1 + Td[1] + Tc1 + Tdf[1] + Ta1) / I[1];
  double al2 = (Tj2 + Td[2] + Tc2 + Tdf[2] + Ta2) / I[2];
  b.acc[0] = a0; b.acc[1] = a1; b.acc[2] = a2;
  b.alp[0] = al0; b.alp[1] = al1; b.alp[2] = al2;

  // ---- _update_motion_states (robot.py:860-875): semi-implicit Euler ----
#pragma unroll
  for (int i = 0; i < 3; i++) { v[i] += b.acc[i] * dt; w[i] += b.alp[i] * dt; }
  double sph, cph, sth, cth;
  sincos(b.eul[0], &sph, &cph);
  sincos(b.eul[1], &sth, &cth);
  double tth = sth / cth;                                      // dynamics.py:21-31
  double er0 = w[0] + sph * tth * w[1] + cph * tth * w[2];
  double er1 = cph * w[1] - sph * w[2];
  double er2 = (sph / cth) * w[1] + (cph / cth) * w[2];
  b.eul[0] += er0 * dt; b.eul[1] += er1 * dt; b.eul[2] += er2 * dt;
  Rot3 R = rotation_zyx(b.eul[0], b.eul[1], b.eul[2]);         // dynamics.py:35-58
  double vw0 = R.r[0] * v[0] + R.r[1] * v[1] + R.r[2] * v[2];
  double vw1 = R.r[3] * v[0] + R.r[4] * v[1] + R.r[5] * v[2];
  double vw2 = R.r[6] * v[0] + R.r[7] * v[1] + R.r[8] * v[2];
  b.vwx = vw0; b.vwy = vw1;
  b.pw[0] += vw0 * dt; b.pw[1] += vw1 * dt; b.pw[2] += vw2 * dt;
#pragma unroll
  for (int i = 0; i < 3; i++) { b.pos[i] += v[i] * dt; b.ang[i] += w[i] * dt; }

  // ---- cycle_time += dt; update_state; update_properties (robot.py:674-678, 640-668) ----
  t = rn::dadd(t, dt);
  b.phase = phase_at(c, t);
  b.prev_volume = b.volume;
  double prev_water_mass = b.prev_volume * p.density;
  double dl = shape_delta(b.phase, t, c.refill, c.T0, (double)c.contraction32, c.contract_rate, c.release_rate);
  b.length = p.init_length - dl;
  b.width = p.init_width + dl;
  refresh_shape_f64(p, b);
  b.mass_rate = (b.water_mass - prev_water_mass) / dt;
  b.com = center_of_mass_x(p, b.length, b.width, b.water_mass);
  b.com_rate = (b.com - b.prev_com) / dt;      // robot.py:901-910
  b.prev_com = b.com;
  b.com_acc = (b.com_rate - b.prev_com_rate) / dt;   // robot.py:912-922
  b.prev_com_rate = b.com_rate;
}


// Robot.step_through_cycle's loop (robot.py:756-757).  Returns K, or -1 if the cycle would run
// past SALP_MAX_SUBSTEPS (impossible for Box actions; guards against non-finite actions).
struct SalpDerived;   // host-derived constants of the mixed loop (salp_loop_mixed.cuh); unused here
// Per-env randomisation context of one cycle (SalpParams.randomization != 0, mixed loop only)
struct RandCtx {
  uint64_t seed;
  int64_t gid;
  uint32_t episode, cycle;
  float ou_fx, ou_fy, ou_tz;      // OUDisturbance.state: force x, y and torque z (robot.py:279-280, 796-838)
};
#define SALP_PRECISION_MIXED_RANDOMIZED 2   // internal: MIXED with SalpParams.randomization != 0
template <int PREC>
SALP_HD int run_cycle(const SalpParams& p, const SalpDerived& dv, const CyclePlan& c, const double* time_table,
                      Body64& b, double& t, RandCtx* rc);

template <>
SALP_HD int run_cycle<SALP_PRECISION_F64>(const SalpParams& p, const SalpDerived& dv, const CyclePlan& c,
                                          const double* time_table, Body64& b, double& t, RandCtx* rc) {
  (void)time_table;
  (void)dv;
  (void)rc;
  refresh_shape_f64(p, b);
  b.mass_rate = (b.water_mass - b.prev_volume * p.density) / p.dt;
  t = 0.0;
  int K = 0;
  while (cycle_running(c, t)) {
    if (K >= SALP_MAX_SUBSTEPS) return -1;
    substep_f64(p, c, b, t);
    K++;
  }
  if (K > 0) b.speed_world = sqrt(b.vwx * b.vwx + b.vwy * b.vwy);
  return K;
}
