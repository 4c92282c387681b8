// salp_device.cuh -- device-side building blocks of the SALP hot path (sm_100a).
//
// Everything here is per-env scalar math that lives in registers.  Reference citations are
// file:line in Avielstein/GRASP_LAB_SALP src/.
#pragma once
#include "salp_common.cuh"

#define SALP_DEV SALP_HD

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

// --------------------------------------------------------------------------------------------
// float32 sin/cos bit-identical to numpy's SIMD float32 kernels (robot.py:76 evaluates
// np.cos/np.sin on an np.float32 yaw; SURVEY.md hard part 1).  Cody-Waite reduction + minimax
// polynomials; the multiply/adds that numpy rounds separately use rn::fmul/rn::fadd so nvcc
// cannot contract them, every fmaf is a real single-rounding FMA.  Valid for |x| <= 71476.
// --------------------------------------------------------------------------------------------
SALP_DEV void np_sincosf(float x, float& sn, float& cs) {
  float q = rn::fadd(rn::fadd(rn::fmul(x, 0x1.45f306p-1f), 0x1.8p+23f), -0x1.8p+23f);
  float r = rn::ffma(q, -0x1.921fb0p+0f, x);
  r = rn::ffma(q, -0x1.5110b4p-22f, r);
  r = rn::ffma(q, -0x1.846988p-48f, r);
  float r2 = rn::fmul(r, r);
  float C = rn::ffma(rn::ffma(rn::ffma(rn::ffma(0x1.98e616p-16f, r2, -0x1.6c06dcp-10f), r2,
                                          0x1.55553cp-5f), r2, -0.5f), r2, 1.0f);
  float S = rn::ffma(rn::ffma(rn::ffma(rn::ffma(rn::ffma(0x1.7d3bbcp-19f, r2, -0x1.a06bbap-13f),
                                                    r2, 0x1.11119ap-7f), r2, -0x1.555556p-3f), r2, 0.0f),
                      r, r);
  int k = (int)q;
  cs = (k & 1) ? S : C;
  sn = (k & 1) ? C : S;
  if ((k + 1) & 2) cs = -cs;
  if (k & 2) sn = -sn;
}

// --------------------------------------------------------------------------------------------
// Philox4x32-10 scene sampler: one counter-based stream per (global env id, episode, draw),
// so results do not depend on how envs are sharded over GPUs.
// --------------------------------------------------------------------------------------------
SALP_DEV void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
// A point uniform in the tank box, cast to float32 like the reference does
// (np.random.uniform(lo, hi) = lo + (hi-lo)*u; salp_robot_env.py:484-487,533,547-550).
SALP_DEV void sample_point(const SalpParams& p, uint64_t seed, int64_t gid, int ep, int draw,
                           float& x, float& y) {
  uint32_t c[4] = {(uint32_t)gid, (uint32_t)((uint64_t)gid >> 32), (uint32_t)ep, (uint32_t)draw};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  double u0 = rn::dmul(rn::dadd((double)c[0], 0.5), 1.0 / 4294967296.0);
  double u1 = rn::dmul(rn::dadd((double)c[1], 0.5), 1.0 / 4294967296.0);
  x = (float)rn::dadd(p.tank_x_min, rn::dmul(rn::dsub(p.tank_x_max, p.tank_x_min), u0));
  y = (float)rn::dadd(p.tank_y_min, rn::dmul(rn::dsub(p.tank_y_max, p.tank_y_min), u1));
}
// Randomisation streams (SalpParams.randomization): one Philox block per (global env id, episode,
// cycle, purpose), keyed differently from the scene sampler.  u01: uniform in (0, 1).
#define SALP_RNG_ACTION 1u
#define SALP_RNG_OBS 2u          // .. 3
#define SALP_RNG_DYNAMICS 4u     // .. 7
#define SALP_RNG_OU 0x1000u      // + substep index
SALP_DEV void rand_block(uint64_t seed, int64_t gid, uint32_t episode, uint32_t cycle, uint32_t purpose,
                         uint32_t out[4]) {
  out[0] = (uint32_t)gid;
  out[1] = (uint32_t)((uint64_t)gid >> 32);
  out[2] = episode;
  out[3] = (cycle << 16) | purpose;
  philox4x32_10(out, (uint32_t)seed ^ 0x5A17C0DEu, (uint32_t)(seed >> 32) ^ 0x9E3779B9u);
}
SALP_DEV float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }
// randomize_scalar_jit (geometry.py:208-222): value * U(1 - u, 1 + u), clipped to [lo, hi]
SALP_DEV float randomize_scalar(float value, float uncertainty, uint32_t bits, float lo, float hi) {
  float s = value * (1.0f + uncertainty * (2.0f * u01(bits) - 1.0f));
  return fminf(fmaxf(s, lo), hi);
}

// ... with the default (NaN) bounds: the clip is min(max(sample, v (1 - u)), v (1 + u)), which for a
// NEGATIVE value has its bounds the wrong way round and always returns v (1 + u) -- a quirk of the
// reference that e.g. observation randomisation inherits (verified on the live reference:
// tests/golden/ref_randstats.npz, samples_observation, columns obs0 / obs4).
SALP_DEV float randomize_scalar_default_bounds(float value, float uncertainty, uint32_t bits) {
  if (value < 0.0f) return value * (1.0f + uncertainty);
  return value * (1.0f + uncertainty * (2.0f * u01(bits) - 1.0f));
}

SALP_DEV float dist2f(float ax, float ay, float bx, float by) {  // float32 norm, no contraction
  float dx = rn::fsub(ax, bx), dy = rn::fsub(ay, by);
  return rn::fsqrt(rn::fadd(rn::fmul(dx, dx), rn::fmul(dy, dy)));
}

// --------------------------------------------------------------------------------------------
// geometry.py, templated on the arithmetic type
// --------------------------------------------------------------------------------------------
template <typename T>
SALP_DEV T ellipsoid_volume(T length, T width) {  // geometry.py:79-81
  T wh = width / T(2);
  return (T(4.0 / 3.0) * T(M_PI)) * (length / T(2)) * (wh * wh);
}
template <typename T>
SALP_DEV void cross_sections(T length, T width, T a[3]) {  // geometry.py:68-75
  T wh = width / T(2), lh = length / T(2);
  a[0] = T(M_PI) * wh * wh;
  a[1] = T(M_PI) * lh * wh;
  a[2] = a[1];
}
// geometry.py:134-183, diagonal.  All placeholder dimensions are 0 in the reference, so only the
// parallel-axis terms of buoy/tube/nozzle and the skin/water ellipsoid terms survive.  The
// density is the literal 1000 of geometry.py:141,168 (not params.density).
template <typename T>
SALP_DEV void inertia_diag(T length, T width, T nozzle_mass, T I[3]) {
  const T mass_buoy = T(0.195), skin_mass = T(0.145), tube_mass = T(0.414);
  const T tube_volume = T(3.14159265358979 * ((0.058 / 2.0) * (0.058 / 2.0)) * 0.15);
  const T net_tube_mass = tube_mass - tube_volume * T(1000.0);
  T lh = length / T(2), wh = width / T(2);
  T lh2 = lh * lh, wh2 = wh * wh;
  T d = lh - T(0.08), e = lh + T(0.025);
  T wme = T(1000.0) * ellipsoid_volume(length, width);
  T skin = T(1.0 / 3.0) * skin_mass;
  T water = T(0.2) * wme;
  I[0] = skin * (wh2 + wh2) + water * (wh2 + wh2);
  T yy = mass_buoy * lh2 + net_tube_mass * (d * d) + skin * (lh2 + wh2) + water * (lh2 + wh2) +
         nozzle_mass * (e * e);
  I[1] = yy;
  I[2] = yy;
}
// geometry.py:187-203 (x component; y and z are identically 0)
template <typename T>
SALP_DEV T center_of_mass_x(const SalpParams& p, T length, T width, T water_mass) {
  T pos_buoy = length / T(2);
  T pos_tube = length / T(2) - T(0.08);
  T pos_nozzle = -length / T(2) - T(0.025) + T(0.05);
  T wme = T(1000) * ellipsoid_volume(length, width);
  T tv = T(1000) * T(p.tube_volume);
  T pos_water = (T(0) - tv * pos_tube) / (wme - tv);
  T total = T(p.tube_mass) + T(p.nozzle_mass) + T(p.buoy_mass) + T(p.skin_mass) + water_mass;
  return (T(p.tube_mass) * pos_tube + T(p.nozzle_mass) * pos_nozzle + T(p.buoy_mass) * pos_buoy +
          water_mass * pos_water) / total;
}
// geometry.py:105-123: returns the interpolation weight; cd_i = hi_i - w*(hi_i - lo_i)
template <typename T>
SALP_DEV T drag_interp_weight(const SalpParams& p, T length, T width) {
  T aspect = length / width;
  T init_aspect = T(p.init_length / p.init_width);
  double cl = p.init_length - p.max_contraction;
  double cw = p.init_length - cl + p.init_width;
  T end_aspect = T(cl / cw);
  T nr = (aspect - end_aspect) / (init_aspect - end_aspect);
  nr = nr < T(0) ? T(0) : nr;
  nr = nr > T(1) ? T(1) : nr;
  return nr;
}
// geometry.py:40-50 and :54-64; `dl` is what is subtracted from init_length / added to init_width
SALP_DEV double shape_delta(int phase, double t, double refill, double T0, double contraction,
                            double contract_rate, double release_rate) {
  // branch-free on purpose (selects): inside the substep loop a branch here would split the basic
  // block and keep the scheduler from overlapping the fp64 shape chain with the fp32 chains
  const double refill_dl = (t < refill) ? t * contract_rate : contraction;
  const double jet_dl = contraction - (t - T0) * release_rate;
  const double d = (phase == 1) ? jet_dl : 0.0;
  return (phase == 0) ? refill_dl : d;
}

// --------------------------------------------------------------------------------------------
// dynamics.py frame helpers (float64; used once per env-step by reward/obs and per substep by
// the reference-mode loop)
// --------------------------------------------------------------------------------------------
struct Rot3 { double r[9]; };
SALP_DEV Rot3 rotation_zyx(double phi, double theta, double psi) {  // dynamics.py:35-57
  double sph, cph, sth, cth, sps, cps;
  sincos(phi, &sph, &cph);
  sincos(theta, &sth, &cth);
  sincos(psi, &sps, &cps);
  Rot3 R;
  R.r[0] = cps * cth; R.r[1] = cps * sth * sph - sps * cph; R.r[2] = cps * sth * cph + sps * sph;
  R.r[3] = sps * cth; R.r[4] = sps * sth * sph + cps * cph; R.r[5] = sps * sth * cph - cps * sph;
  R.r[6] = -sth;      R.r[7] = cth * sph;                   R.r[8] = cth * cph;
  return R;
}
SALP_DEV void to_body_frame_xy(const Rot3& R, double x, double y, double& bx, double& by) {
  // dynamics.py:61-84 with vector (x, y, 0): R^T v, first two components
  bx = R.r[0] * x + R.r[3] * y;
  by = R.r[1] * x + R.r[4] * y;
}

// --------------------------------------------------------------------------------------------
// Per-cycle plan: everything Nozzle.set_yaw_angle/solve_angles, Robot.set_control and the head
// of Robot.step_through_cycle compute before the substep loop (robot.py:62-98, 544-592, 742).
// All float32/float64 typing quirks of SURVEY.md hard parts 1-2 live here.
// --------------------------------------------------------------------------------------------
struct CyclePlan {
  float contraction32, coast32, yaw32;
  double angle1, angle2, turn;
  double refill, jet, T0, Tjet;
  double contract_rate, release_rate;   // values of np.float32 quotients (or 0.0)
  bool total_is_f32;
  float total32;
  double total64;       // == (double)total32 when total_is_f32
  double dir[3];        // Nozzle.get_nozzle_direction(), constant within the cycle
};

SALP_DEV double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

SALP_DEV CyclePlan make_cycle_plan(const SalpParams& p, float a0, float a1, float a2,
                                   double prev_angle1, double prev_angle2, const uint32_t* action_noise = nullptr) {
  CyclePlan c;
  // _rescale_action (salp_robot_env.py:166-174): float32 products under NEP 50
  c.contraction32 = rn::fmul(a0, (float)0.06);
  c.coast32 = rn::fmul(a1, (float)10.0);
  c.yaw32 = rn::fmul(a2, (float)(M_PI / 2));
  if (action_noise) {   // _randomize_actions (salp_robot_env.py:176-181), on the rescaled action
    c.contraction32 = randomize_scalar(c.contraction32, 0.1f, action_noise[0], 0.0f, 1.0f);
    c.coast32 = randomize_scalar(c.coast32, 0.1f, action_noise[1], 0.0f, 20.0f);
    c.yaw32 = randomize_scalar(c.yaw32, 0.1f, action_noise[2], -(float)(M_PI / 2), (float)(M_PI / 2));
  }
  // Nozzle.solve_angles (robot.py:71-98)
  float s32, c32;
  np_sincosf(c.yaw32, s32, c32);
  double ty = -(double)s32, tz = (double)c32;     // R_br^T @ -[cos, sin, 0] = [0, -sin, cos]
  double val2 = clampd(rn::dsub(rn::dmul(2.0, tz), 1.0), -1.0, 1.0);
  double angle2 = acos(val2);
  double angle1;
  double sa2, ca2;
  sincos(angle2, &sa2, &ca2);
  double a = 0.5 * (ca2 - 1.0);
  double b = sqrt(2.0) * sa2 / 2.0;
  if (angle2 == 0.0) {
    angle1 = 0.0;
  } else {
    double val1 = clampd(ty / sqrt(rn::dadd(rn::dmul(a, a), rn::dmul(b, b))), -1.0, 1.0);
    angle1 = asin(val1) - atan2(b, a);
  }
  if (angle1 <= -M_PI) angle1 += 2 * M_PI;
  else if (angle1 > M_PI) angle1 -= 2 * M_PI;
  c.angle1 = angle1;
  c.angle2 = angle2;
  // Nozzle.set_angles -> _nozzle_turn_time (robot.py:173-185)
  c.turn = rn::dadd(rn::ddiv(fabs(rn::dsub(angle1, prev_angle1)), p.nozzle_angle_speed),
                     rn::ddiv(fabs(rn::dsub(angle2, prev_angle2)), p.nozzle_angle_speed));
  // Nozzle.get_nozzle_direction (robot.py:138-150): R_br Rz(angle1) Rfix(gamma) Rz(angle2) [cg,0,sg]
  {
    double sg, cg, s1, c1;
    sincos(p.nozzle_gamma, &sg, &cg);
    sincos(angle1, &s1, &c1);
    double u0 = ca2 * cg, u1 = sa2 * cg, u2 = sg;
    double q0 = cg * u0 - sg * u2, q1 = u1, q2 = sg * u0 + cg * u2;
    double r0 = c1 * q0 - s1 * q1, r1 = s1 * q0 + c1 * q1, r2 = q2;
    c.dir[0] = -r2; c.dir[1] = r1; c.dir[2] = r0;
  }
  // Robot.set_control (robot.py:589-592): numba float32 specialisation squares in float32
  float sq = rn::fmul(c.contraction32, c.contraction32);
  double x = (double)c.contraction32;
  c.refill = rn::dadd(rn::dadd(rn::dmul(p.refill_poly[0], (double)sq), rn::dmul(p.refill_poly[1], x)),
                       p.refill_poly[2]);
  c.jet = rn::dadd(rn::dadd(rn::dmul(p.jet_poly[0], (double)sq), rn::dmul(p.jet_poly[1], x)),
                    p.jet_poly[2]);
  c.contract_rate = c.refill > 0.0 ? (double)rn::fdiv(c.contraction32, (float)c.refill) : 0.0;
  c.release_rate = c.jet > 0.0 ? (double)rn::fdiv(c.contraction32, (float)c.jet) : 0.0;
  c.T0 = fmax(c.refill, c.turn);
  c.Tjet = rn::dadd(c.T0, c.jet);
  // total_cycle_time (robot.py:742): python max() keeps its first argument unless the second is
  // greater; pyfloat + np.float32 is a float32 add, np.float64 + np.float32 a float64 add.
  if (c.turn > c.refill) {
    c.total_is_f32 = false;
    c.total64 = rn::dadd(rn::dadd(c.turn, c.jet), (double)c.coast32);
    c.total32 = 0.0f;
  } else {
    c.total_is_f32 = true;
    c.total32 = rn::fadd((float)rn::dadd(c.refill, c.jet), c.coast32);
    c.total64 = (double)c.total32;
  }
  return c;
}

// `while self.cycle_time < total_cycle_time` (robot.py:756) in the dtype numpy compares in
SALP_DEV bool cycle_running(const CyclePlan& c, double t) {
  return c.total_is_f32 ? ((float)t < c.total32) : (t < c.total64);
}
// Robot.update_state (robot.py:640-649)
SALP_DEV int phase_at(const CyclePlan& c, double t) {
  if (t <= c.T0) return 0;
  if (t <= c.Tjet) return 1;
  bool le = c.total_is_f32 ? ((float)t <= c.total32) : (t <= c.total64);
  return le ? 2 : 3;
}

// K = number of substeps `while cycle_time < total` (robot.py:756) will run: the first k with
// !(t_k < total), t_k = k-fold repeated `+= dt` (the table).  -1 if beyond SALP_MAX_SUBSTEPS
// (or the total is NaN/inf: non-finite action).
SALP_DEV int plan_substeps(const CyclePlan& c, const double* time_table) {
  if (!(c.total64 == c.total64)) return 0;               // NaN total: `t < nan` is False, the loop never runs
  if (cycle_running(c, time_table[SALP_MAX_SUBSTEPS])) return -1;
  // `running(t_k)` is monotone in k (the table increases, also after rounding to float32), and
  // total / dt is within an entry or two of the answer: read a window of five entries around that
  // guess -- five INDEPENDENT loads, one memory latency -- and count the running ones.  (Round 1
  // bisected: 13 dependent reads; round 2 first walked entry by entry from the guess: ~10.)
  const double g = c.total64 * 100.0;                    // 1 / dt = 100 for the reference's dt; any guess is correct, only slower
  const int kg = g < 2.0 ? 0 : (g > (double)(SALP_MAX_SUBSTEPS - 2) ? SALP_MAX_SUBSTEPS - 4 : (int)g - 2);
  {
    const bool r0 = cycle_running(c, time_table[kg]), r1 = cycle_running(c, time_table[kg + 1]),
               r2 = cycle_running(c, time_table[kg + 2]), r3 = cycle_running(c, time_table[kg + 3]),
               r4 = cycle_running(c, time_table[kg + 4]);
    if (r0 && !r4) return kg + (int)r0 + (int)r1 + (int)r2 + (int)r3;      // first entry that is not running
  }
  int lo = 0, hi = SALP_MAX_SUBSTEPS;                     // running(t_k) for k < lo; !running(t_hi)
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (cycle_running(c, time_table[mid])) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// first k in [0, SALP_MAX_SUBSTEPS] with !(t_k < x) (strict) or !(t_k <= x) (non-strict); the
// guess x/dt is within one or two entries of the answer, so this is a couple of table reads.
template <bool STRICT>
SALP_DEV int first_k_past(const double* table, double x, double inv_dt) {
  if (!(x == x)) return 0;                                         // NaN: every comparison is False
  double g = x * inv_dt;
  int k = g < 0.0 ? 0 : (g > (double)SALP_MAX_SUBSTEPS ? SALP_MAX_SUBSTEPS : (int)g);
  {   // window of five independent reads around the guess (one latency); the walk below only in odd cases
    const int k0 = k < 2 ? 0 : (k > SALP_MAX_SUBSTEPS - 2 ? SALP_MAX_SUBSTEPS - 4 : k - 2);
    const double t0 = table[k0], t1 = table[k0 + 1], t2 = table[k0 + 2], t3 = table[k0 + 3], t4 = table[k0 + 4];
    const bool p0 = STRICT ? t0 < x : t0 <= x, p1 = STRICT ? t1 < x : t1 <= x, p2 = STRICT ? t2 < x : t2 <= x,
               p3 = STRICT ? t3 < x : t3 <= x, p4 = STRICT ? t4 < x : t4 <= x;
    if ((p0 || k0 == 0) && !p4) return k0 + (int)p0 + (int)p1 + (int)p2 + (int)p3;
  }
  while (k > 0 && !(STRICT ? table[k - 1] < x : table[k - 1] <= x)) k--;
  while (k < SALP_MAX_SUBSTEPS && (STRICT ? table[k] < x : table[k] <= x)) k++;
  return k;
}


// Integer phase plan of one cycle (robot.py:640-649 evaluated on the table t_j; update j follows
// substep j-1):  phase_j = 0 for j < k_T0, 1 for k_T0 <= j < k_jet, 2/3 afterwards.  The body
// shape moves at updates j <= k_ref (refill ramp and its end) and k_T0 <= j <= k_jet (jet and its
// end); two more updates flush the first/second backward differences.
struct PhasePlan {
  int k_ref, k_T0, k_jet;
  int upd_a_end, upd_b_begin, upd_b_end;
};
SALP_DEV PhasePlan make_phase_plan(const CyclePlan& c, const double* time_table, double inv_dt) {
  PhasePlan pp;
  pp.k_ref = first_k_past<true>(time_table, c.refill, inv_dt);
  pp.k_T0 = first_k_past<false>(time_table, c.T0, inv_dt);
  pp.k_jet = first_k_past<false>(time_table, c.Tjet, inv_dt);
  pp.upd_a_end = (pp.k_ref > 1 ? pp.k_ref : 1) + 2;
  pp.upd_b_begin = pp.k_T0;
  pp.upd_b_end = (pp.k_jet > pp.k_T0 ? pp.k_jet : pp.k_T0) + 2;
  return pp;
}

// Sort key of the K-sort: coarse K bucket (32 substeps) first, then the end of the shape motion
// (8-substep bins, capped), so that the lanes of a warp both finish together AND leave the
// shape-update window together.
SALP_DEV int sort_key(int K, const PhasePlan& pp) {
  int e = pp.upd_b_end < K ? pp.upd_b_end : K;
  e = e >> 3;
  e = e < SALP_SORT_SHAPE_BINS - 1 ? e : SALP_SORT_SHAPE_BINS - 1;
  return (K >> 5) * SALP_SORT_SHAPE_BINS + e;
}
