// salp_env.cuh -- the per-env step and reset bodies (SalpRobotEnv.step / reset, batched).
//
// One thread owns one env.  env_step() is the whole of salp_robot_env.py:196-299 for that env:
// action rescale -> nozzle inverse kinematics -> Robot.set_control -> the K-substep breathing
// cycle (run_cycle<PREC>, state in registers) -> reward, observation, collision, termination,
// episode bookkeeping -> optional SB3-style auto-reset.  Reference citations are file:line in
// Avielstein/GRASP_LAB_SALP src/.
#pragma once
#include "salp_loop_f64.cuh"
#include "salp_loop_mixed.cuh"

// state accessors of env i (record-per-env layout, see salp_common.cuh)
struct Cols {
  const SalpView& v;
  int64_t i;
#if SALP_STATE_AOS
  SALP_HD double& d(int f) const { return v.f64[i * SALP_NUM_F64_FIELDS + f]; }
  SALP_HD float& f(int f) const { return v.f32[i * SALP_NUM_F32 + (f - SALP_F32_BASE)]; }
  SALP_HD int32_t& n(int f) const { return v.i32[i * SALP_NUM_I32 + (f - SALP_I32_BASE)]; }
#else
  SALP_HD double& d(int f) const { return v.f64[(int64_t)f * v.n + i]; }
  SALP_HD float& f(int f) const { return v.f32[(int64_t)(f - SALP_F32_BASE) * v.n + i]; }
  SALP_HD int32_t& n(int f) const { return v.i32[(int64_t)(f - SALP_I32_BASE) * v.n + i]; }
#endif
};

SALP_HD double norm2d(double x, double y) { return sqrt(x * x + y * y); }

SALP_HD void raise_status(const SalpView& v, int code) {
#ifdef __CUDA_ARCH__
  atomicMin(v.status, code);
#else
  if (code < *v.status) *v.status = code;
#endif
}

// _get_observation (salp_robot_env.py:651-670)
SALP_HD void write_observation(const SalpParams& p, const Cols& c, const double pw[3],
                               const double eul[3], const double v[3], double wz, float* obs) {
  double tx = (double)c.f(SALP_F_TARGET_X), ty = (double)c.f(SALP_F_TARGET_Y);
  Rot3 R = rotation_zyx(eul[0], eul[1], eul[2]);
  double bx, by;
  to_body_frame_xy(R, tx - pw[0], ty - pw[1], bx, by);
  obs[0] = (float)bx;
  obs[1] = (float)by;
  obs[2] = (float)v[0];
  obs[3] = (float)v[1];
  obs[4] = (float)wz;
  obs[5] = (float)atan2(by, bx);
  for (int k = 0; k < p.num_obstacles; k++) {
    obs[6 + 2 * k] = (float)((double)c.f(SALP_F_OBSTACLE0_X + 2 * k) - pw[0]);
    obs[7 + 2 * k] = (float)((double)c.f(SALP_F_OBSTACLE0_X + 2 * k + 1) - pw[1]);
  }
}

// generate_target_point("random") + _generate_obstacles (salp_robot_env.py:449-559), or the
// caller's scene pool (salp_set_scene_pool) when one is installed.
SALP_HD void sample_scene(const SalpParams& p, const SalpView& v, const Cols& c, int episode) {
  const int nobs = p.num_obstacles;
  if (v.pool_P > 0) {
    int64_t s = (int64_t)episode % v.pool_P;
    const float* t = v.pool_targets + (c.i * v.pool_P + s) * 2;
    c.f(SALP_F_TARGET_X) = t[0];
    c.f(SALP_F_TARGET_Y) = t[1];
    const float* ob = v.pool_obstacles + (c.i * v.pool_P + s) * nobs * 2;
    for (int k = 0; k < 2 * nobs; k++) c.f(SALP_F_OBSTACLE0_X + k) = ob[k];
    return;
  }
  int64_t gid = v.env_id_offset + c.i;
  int draw = 0;
  float tx, ty;
  sample_point(p, v.seed, gid, episode, draw++, tx, ty);
  c.f(SALP_F_TARGET_X) = tx;
  c.f(SALP_F_TARGET_Y) = ty;
  const float min_clear = 0.5f;                                  // salp_robot_env.py:540-541
  const float min_sep = (float)(2 * p.obstacle_radius + 0.1);    // :542
  for (int k = 0; k < nobs; k++) {
    float px = 0.f, py = 0.f;
    for (int attempt = 0; attempt < 200; attempt++) {
      sample_point(p, v.seed, gid, episode, draw++, px, py);
      bool too_close = false;
      for (int j = 0; j < k; j++)
        too_close |= dist2f(px, py, c.f(SALP_F_OBSTACLE0_X + 2 * j), c.f(SALP_F_OBSTACLE0_X + 2 * j + 1)) < min_sep;
      if (dist2f(px, py, 0.f, 0.f) > min_clear && dist2f(px, py, tx, ty) > min_clear && !too_close) break;
      // after 200 rejections the reference silently places fewer obstacles (shorter obs); this
      // restatement keeps the last candidate instead (probability ~1e-140)
    }
    c.f(SALP_F_OBSTACLE0_X + 2 * k) = px;
    c.f(SALP_F_OBSTACLE0_X + 2 * k + 1) = py;
  }
}

// Robot.__init__ + Nozzle.__init__ + make_env's set_angles(0, 0) (robot.py:20-47, 261-374;
// train_robot.py:11-21): the state of an env that has never been reset.
SALP_HD void env_init(const SalpParams& p, const SalpView& v, int64_t i) {
  Cols c{v, i};
  for (int f = 0; f < SALP_NUM_F64_FIELDS; f++) c.d(f) = 0.0;
  for (int f = SALP_F32_BASE; f < SALP_F32_END; f++) c.f(f) = 0.f;
  for (int f = SALP_I32_BASE; f < SALP_I32_END; f++) c.n(f) = 0;
  c.n(SALP_F_PHASE) = 3;
  double l = p.init_length, w = p.init_width;
  c.d(SALP_F_LENGTH) = l;
  c.d(SALP_F_WIDTH) = w;
  double vol = ellipsoid_volume(l, w) - p.tube_volume;
  c.d(SALP_F_PREV_VOLUME) = vol;
  double I[3];
  inertia_diag(l, w, p.nozzle_mass, I);
  c.d(SALP_F_PREV_I_X) = I[0];
  c.d(SALP_F_PREV_I_Y) = I[1];
  c.d(SALP_F_PREV_I_Z) = I[2];
  double com = center_of_mass_x(p, l, w, p.density * vol);
  c.d(SALP_F_COM_X) = com;
  c.d(SALP_F_PREV_COM_X) = com;
}

// SalpRobotEnv.reset (salp_robot_env.py:114-155) incl. Robot.reset (robot.py:452-501).
// Works on the HBM columns directly (a reset is ~100 B of traffic and ~200 flop).
SALP_HD void env_reset(const SalpParams& p, const SalpView& v, int64_t i, float* obs) {
  Cols c{v, i};
  int episode = c.n(SALP_F_EPISODE_INDEX);
  sample_scene(p, v, c, episode);
  c.n(SALP_F_EPISODE_INDEX) = episode + 1;
  // Robot.reset: motion state to zero; the nozzle (angle1, angle2, yaw) is NOT touched
  for (int f = SALP_F_VEL_X; f <= SALP_F_PREVANGLE_Z; f++) c.d(f) = 0.0;
  c.d(SALP_F_SPEED_WORLD) = 0.0;
  c.d(SALP_F_OU_FORCE_X) = 0.0;        // force_disturbance.reset(), torque_disturbance.reset() (robot.py:454-455)
  c.d(SALP_F_OU_FORCE_Y) = 0.0;
  c.d(SALP_F_OU_TORQUE_Z) = 0.0;
  c.n(SALP_F_CYCLE) = 0;
  c.n(SALP_F_PHASE) = 3;
  // centre of mass is evaluated BEFORE length/width are restored (robot.py:478 vs :485-486)
  {
    double l = c.d(SALP_F_LENGTH), w = c.d(SALP_F_WIDTH);
    double wm = p.density * (ellipsoid_volume(l, w) - p.tube_volume);
    double com = center_of_mass_x(p, l, w, wm);
    c.d(SALP_F_COM_X) = com;
    c.d(SALP_F_PREV_COM_X) = com;
    c.d(SALP_F_COM_RATE_X) = 0.0;        // (com - prev_com)/dt with prev_com == com
    c.d(SALP_F_PREV_COM_RATE_X) = 0.0;
    c.d(SALP_F_COM_ACC_X) = 0.0;
  }
  double l0 = p.init_length, w0 = p.init_width;
  c.d(SALP_F_LENGTH) = l0;
  c.d(SALP_F_WIDTH) = w0;
  c.d(SALP_F_PREV_VOLUME) = ellipsoid_volume(l0, w0) - p.tube_volume;
  double I[3];
  inertia_diag(l0, w0, p.nozzle_mass, I);
  c.d(SALP_F_PREV_I_X) = I[0];
  c.d(SALP_F_PREV_I_Y) = I[1];
  c.d(SALP_F_PREV_I_Z) = I[2];
  // env part (salp_robot_env.py:127-153)
  double tx = (double)c.f(SALP_F_TARGET_X), ty = (double)c.f(SALP_F_TARGET_Y);
  double d0 = norm2d(0.0 - tx, 0.0 - ty);
  c.d(SALP_F_PREV_DIST) = d0;
  c.f(SALP_F_PREV_ACTION0) = 0.f;
  c.f(SALP_F_PREV_ACTION1) = 0.f;
  c.f(SALP_F_PREV_ACTION2) = 0.f;
  c.n(SALP_F_EP_LENGTH) = 0;
  for (int f = SALP_F_EP_RETURN; f <= SALP_F_EP_SUBSTEPS; f++) c.d(f) = 0.0;
  c.d(SALP_F_EP_INITIAL_DISTANCE) = d0;
  if (obs) {
    const double z3[3] = {0.0, 0.0, 0.0};
    write_observation(p, c, z3, z3, z3, 0.0, obs);
  }
}

// _calculate_episode_metrics (salp_robot_env.py:399-447) + SB3 Monitor's r / l
SALP_HD void write_episode_metrics(const Cols& c, double last_x, double last_y, double final_dist,
                                   int ep_len, double* m) {
  for (int k = 0; k < SALP_NUM_EPISODE_METRICS; k++) m[k] = 0.0;
  double path = c.d(SALP_F_EP_PATH_LENGTH);
  m[SALP_EM_RETURN] = c.d(SALP_F_EP_RETURN);
  m[SALP_EM_LENGTH] = (double)ep_len;
  m[SALP_EM_PATH_LENGTH] = path;
  double direct = norm2d(last_x, last_y);
  m[SALP_EM_DIRECT_DISTANCE] = direct;
  m[SALP_EM_PATH_EFFICIENCY] = path > 0 ? direct / path : 0.0;
  m[SALP_EM_FINAL_DISTANCE] = final_dist;
  m[SALP_EM_INITIAL_DISTANCE] = c.d(SALP_F_EP_INITIAL_DISTANCE);
  double n = (double)ep_len;
  if (ep_len > 0) {
    m[SALP_EM_AVG_COMPRESSION] = c.d(SALP_F_EP_SUM_A0) / n;
    m[SALP_EM_AVG_COAST_TIME] = c.d(SALP_F_EP_SUM_A1) / n;
    m[SALP_EM_AVG_NOZZLE_ANGLE] = c.d(SALP_F_EP_SUM_ABS_A2) / n;
    for (int k = 0; k < 7; k++) m[SALP_EM_AVG_REWARD_TRACK + k] = c.d(SALP_F_EP_SUM_TERM0 + k) / n;
  }
  m[SALP_EM_AVG_VELOCITY] = c.d(SALP_F_EP_SUM_SPEED) / (n + 1.0);
  m[SALP_EM_TOTAL_SUBSTEPS] = c.d(SALP_F_EP_SUBSTEPS);
}

SALP_HD void load_body(const Cols& c, Body64& b) {
#pragma unroll
  for (int k = 0; k < 3; k++) {
    b.v[k] = c.d(SALP_F_VEL_X + k);
    b.w[k] = c.d(SALP_F_ANGVEL_X + k);
    b.eul[k] = c.d(SALP_F_EULER_X + k);
    b.pw[k] = c.d(SALP_F_POSW_X + k);
    b.acc[k] = c.d(SALP_F_ACC_X + k);
    b.alp[k] = c.d(SALP_F_ANGACC_X + k);
    b.pos[k] = c.d(SALP_F_POS_X + k);
    b.ang[k] = c.d(SALP_F_ANGLE_X + k);
    b.prevI[k] = c.d(SALP_F_PREV_I_X + k);
  }
  b.phase = c.n(SALP_F_PHASE);
  b.length = c.d(SALP_F_LENGTH);
  b.width = c.d(SALP_F_WIDTH);
  b.prev_volume = c.d(SALP_F_PREV_VOLUME);
  b.com = c.d(SALP_F_COM_X);
  b.prev_com = c.d(SALP_F_PREV_COM_X);
  b.com_rate = c.d(SALP_F_COM_RATE_X);
  b.prev_com_rate = c.d(SALP_F_PREV_COM_RATE_X);
  b.com_acc = c.d(SALP_F_COM_ACC_X);
  b.speed_world = c.d(SALP_F_SPEED_WORLD);
}

SALP_HD void store_body(const Cols& c, const Body64& b) {
#pragma unroll
  for (int k = 0; k < 3; k++) {
    c.d(SALP_F_VEL_X + k) = b.v[k];
    c.d(SALP_F_ANGVEL_X + k) = b.w[k];
    c.d(SALP_F_EULER_X + k) = b.eul[k];
    c.d(SALP_F_POSW_X + k) = b.pw[k];
    c.d(SALP_F_ACC_X + k) = b.acc[k];
    c.d(SALP_F_ANGACC_X + k) = b.alp[k];
    c.d(SALP_F_POS_X + k) = b.pos[k];
    c.d(SALP_F_ANGLE_X + k) = b.ang[k];
    c.d(SALP_F_PREV_I_X + k) = b.prevI[k];
  }
  c.n(SALP_F_PHASE) = b.phase;
  c.d(SALP_F_LENGTH) = b.length;
  c.d(SALP_F_WIDTH) = b.width;
  c.d(SALP_F_PREV_VOLUME) = b.prev_volume;
  c.d(SALP_F_COM_X) = b.com;
  c.d(SALP_F_PREV_COM_X) = b.prev_com;
  c.d(SALP_F_COM_RATE_X) = b.com_rate;
  c.d(SALP_F_PREV_COM_RATE_X) = b.prev_com_rate;
  c.d(SALP_F_COM_ACC_X) = b.com_acc;
  c.d(SALP_F_SPEED_WORLD) = b.speed_world;
}

// ---- SalpRobotEnv.step (salp_robot_env.py:196-299) for env i, in three parts so that the fused
// kernels (one thread does everything) and the warp-specialised pipeline kernel (three warps
// share one env group, salp_pipe_kernel.cuh) run the same code:
//   env_step_begin : reads only  -- action rescale, nozzle IK, set_control, state load
//   run_cycle<PREC>              -- the K-substep loop (or the pipeline)
//   env_step_end   : all writes  -- state store, reward, obs, termination, bookkeeping, auto-reset
struct StepCtx {
  float a0, a1, a2;
  CyclePlan plan;
  double avg_vy, avg_wz;      // lag-by-one cycle averages (robot.py:744-745)
  double last_x, last_y;      // episode_positions[-1]
  int cycle;
  RandCtx rc;                 // only meaningful when SalpParams.randomization != 0
};

// The epilogue (env_step_end) reads the tail of the env's record -- episode accumulators, target,
// obstacles -- which the prologue does not touch: after a long substep loop (or a flushed L2) those
// lines come from DRAM on the critical path.  Ask for them now.
SALP_HD void prefetch_record_tail(const SalpView& v, int64_t i) {
#if defined(__CUDA_ARCH__) && SALP_STATE_AOS
  const char* f64 = reinterpret_cast<const char*>(v.f64 + i * SALP_NUM_F64_FIELDS);
  asm volatile("prefetch.global.L2 [%0];" ::"l"(f64 + 8 * SALP_F_PREV_DIST));
  asm volatile("prefetch.global.L2 [%0];" ::"l"(f64 + 8 * SALP_F_EP_SUM_TERM3));
  asm volatile("prefetch.global.L2 [%0];" ::"l"(f64 + 8 * (SALP_NUM_F64_FIELDS - 1)));
  asm volatile("prefetch.global.L2 [%0];" ::"l"(v.f32 + i * SALP_NUM_F32));
  asm volatile("prefetch.global.L2 [%0];" ::"l"(v.f32 + i * SALP_NUM_F32 + SALP_NUM_F32 - 1));
#else
  (void)v; (void)i;
#endif
}

SALP_HD void env_step_begin(const SalpParams& p, const SalpView& v, const SalpStepIO& io, int64_t i, StepCtx& cx,
                            Body64& b) {
  Cols c{v, i};
  prefetch_record_tail(v, i);
  cx.a0 = io.actions[3 * i];
  cx.a1 = io.actions[3 * i + 1];
  cx.a2 = io.actions[3 * i + 2];
  // :201-209  rescale, Nozzle.set_yaw_angle / solve_angles, Robot.set_control
  cx.cycle = c.n(SALP_F_CYCLE) + 1;
  if (p.randomization) {
    cx.rc.seed = v.seed;
    cx.rc.gid = v.env_id_offset + i;
    cx.rc.episode = (uint32_t)c.n(SALP_F_EPISODE_INDEX);
    cx.rc.cycle = (uint32_t)cx.cycle;
    cx.rc.ou_fx = (float)c.d(SALP_F_OU_FORCE_X);
    cx.rc.ou_fy = (float)c.d(SALP_F_OU_FORCE_Y);
    cx.rc.ou_tz = (float)c.d(SALP_F_OU_TORQUE_Z);
  }
  if (p.randomization & SALP_RAND_ACTION) {
    uint32_t r[4];
    rand_block(cx.rc.seed, cx.rc.gid, cx.rc.episode, cx.rc.cycle, SALP_RNG_ACTION, r);
    cx.plan = make_cycle_plan(p, cx.a0, cx.a1, cx.a2, c.d(SALP_F_NOZZLE_ANGLE1), c.d(SALP_F_NOZZLE_ANGLE2), r);
  } else {
    cx.plan = make_cycle_plan(p, cx.a0, cx.a1, cx.a2, c.d(SALP_F_NOZZLE_ANGLE1), c.d(SALP_F_NOZZLE_ANGLE2));
  }
  // :210  Robot.step_through_cycle (robot.py:740-757)
  load_body(c, b);
  const double tot = cx.plan.total64;
  // lag-by-one: displacement of the PREVIOUS cycle over the NEW total (robot.py:744-748)
  cx.avg_vy = (b.pos[1] - c.d(SALP_F_PREVPOS_Y)) / tot;
  cx.avg_wz = (b.ang[2] - c.d(SALP_F_PREVANGLE_Z)) / tot;
  cx.last_x = b.pw[0];
  cx.last_y = b.pw[1];
}

// `pos0` / `ang0`: body-frame integrals at the START of the cycle (they become prev_position /
// prev_angle, robot.py:747-748); K, t: substeps run and cycle_time reached.
//
// `obs_row` / `tobs_row`: where this env's observation and terminal observation are assembled.  The
// default is their final place in io.obs / io.terminal_obs; the step kernel passes rows of a
// shared-memory tile instead and writes the tile out warp-wide afterwards -- coalesced, and
// without ever READING the output arrays, which may then be mapped host memory (salp_step_host).
SALP_HD void env_step_end(const SalpParams& p, const SalpView& v, const SalpStepIO& io, uint32_t flags, int64_t i,
                          const StepCtx& cx, const double pos0[3], const double ang0[3], Body64& b, int K, double t,
                          float* obs_row = nullptr, float* tobs_row = nullptr) {
  Cols c{v, i};
  const int D = SALP_OBS_BASE + 2 * p.num_obstacles;
  if (!obs_row) {
    obs_row = io.obs + i * D;
    tobs_row = io.terminal_obs ? io.terminal_obs + i * D : nullptr;
  }
  const CyclePlan& plan = cx.plan;
  const float a0 = cx.a0, a1 = cx.a1, a2 = cx.a2;
  const double avg_vy = cx.avg_vy, avg_wz = cx.avg_wz, last_x = cx.last_x, last_y = cx.last_y;
  const int cycle = cx.cycle;
  c.f(SALP_F_NOZZLE_YAW) = plan.yaw32;
  c.d(SALP_F_NOZZLE_ANGLE1) = plan.angle1;
  c.d(SALP_F_NOZZLE_ANGLE2) = plan.angle2;
  // enable_latency (salp_robot_env.py:294-297): the extra set_control() after the step only bumps
  // Robot.cycle (contraction 0, nothing is integrated) -- the timeout test below still sees `cycle`
  c.n(SALP_F_CYCLE) = (p.randomization & SALP_RAND_LATENCY) ? cycle + 1 : cycle;
  if (p.randomization & SALP_RAND_DISTURBANCE) {
    c.d(SALP_F_OU_FORCE_X) = (double)cx.rc.ou_fx;
    c.d(SALP_F_OU_FORCE_Y) = (double)cx.rc.ou_fy;
    c.d(SALP_F_OU_TORQUE_Z) = (double)cx.rc.ou_tz;
  }
#pragma unroll
  for (int k = 0; k < 3; k++) {
    c.d(SALP_F_PREVPOS_X + k) = pos0[k];
    c.d(SALP_F_PREVANGLE_X + k) = ang0[k];
  }
  if (K < 0) raise_status(v, SALP_ERR_RANGE);
  store_body(c, b);

  // Where the reference RAISES: a cycle whose contraction makes jet_time a fraction of one substep
  // (a0 in about [0.0905, 0.094]) drives the semi-implicit Euler integrator unstable, the state
  // overflows, and np.linalg.solve inside dynamics.py:6-10 throws LinAlgError("Array must not
  // contain infs or NaNs") -- the reference process dies.  A batch cannot raise per env: the env
  // is truncated instead (reward = -out_of_bounds_penalty, observation sanitised to finite
  // values, episode metric SALP_EM_NONFINITE = 1) as soon as its state is non-finite at the end
  // of a cycle, which is at most one env-step earlier than the reference's exception.
  const bool crashed = !(isfinite(b.v[0]) && isfinite(b.v[1]) && isfinite(b.v[2]) && isfinite(b.w[0]) &&
                         isfinite(b.w[1]) && isfinite(b.w[2]) && isfinite(b.pw[0]) && isfinite(b.pw[1]) &&
                         isfinite(b.pw[2]) && isfinite(b.eul[0]) && isfinite(b.eul[1]) && isfinite(b.eul[2]));

  // :238-243
  const double path = c.d(SALP_F_EP_PATH_LENGTH) + norm2d(b.pw[0] - last_x, b.pw[1] - last_y);
  c.d(SALP_F_EP_PATH_LENGTH) = path;
  c.d(SALP_F_EP_SUM_SPEED) += b.speed_world;
  const double tx = (double)c.f(SALP_F_TARGET_X), ty = (double)c.f(SALP_F_TARGET_Y);
  const double dx = b.pw[0] - tx, dy = b.pw[1] - ty;
  const double dist = norm2d(dx, dy);

  // :246  _calculate_reward_with_components (:349-397)
  double terms[7];
  terms[0] = (-dist + c.d(SALP_F_PREV_DIST)) * 100;
  c.d(SALP_F_PREV_DIST) = dist;
  // The body-frame vector of the reward (pos - target, :357-359) is the exact negative of the one in
  // the observation (target - pos, :653-656), and r_heading takes atan2 of ITS negative: both use
  // the same atan2(by, bx) of the body-frame target vector.  One rotation, one atan2.
  Rot3 R = rotation_zyx(b.eul[0], b.eul[1], b.eul[2]);
  double bx, by;
  to_body_frame_xy(R, -dx, -dy, bx, by);
  const double heading = atan2(by, bx);
  terms[1] = -0.5 * fabs(heading);
  const int ep_len = c.n(SALP_F_EP_LENGTH);
  if (ep_len == 0) {
    // first step of an episode: prev_action is reset()'s float64 zeros (:128) -> float64 arithmetic
    double ch = (double)a2 - 0.0;
    terms[2] = -1.0 * (ch * ch);
  } else {
    float ch = rn::fsub(a2, c.f(SALP_F_PREV_ACTION2));
    terms[2] = (double)rn::fmul(-1.0f, rn::fmul(ch, ch));
  }
  terms[3] = -10.0 * fabs(avg_wz);
  terms[4] = -0.1;
  terms[5] = -100.0 * fabs(avg_vy);
  terms[6] = 0.0;
  // :255 _check_obstacle_collision (:561-568) shares the distances with the proximity penalty
  const double cur_len = p.init_length - shape_delta(b.phase, t, plan.refill, plan.T0, (double)plan.contraction32,
                                                     plan.contract_rate, plan.release_rate);
  bool hit = false;
  if (p.num_obstacles > 0) {
    double min_dist = INFINITY;
    for (int k = 0; k < p.num_obstacles; k++) {
      double d = norm2d(b.pw[0] - (double)c.f(SALP_F_OBSTACLE0_X + 2 * k),
                        b.pw[1] - (double)c.f(SALP_F_OBSTACLE0_X + 2 * k + 1));
      min_dist = d < min_dist ? d : min_dist;
      hit |= d < p.obstacle_radius + cur_len / 2;
    }
    const double danger = 2.0 * p.obstacle_radius;
    if (min_dist < danger) terms[6] = -1.0 * (1.0 - min_dist / danger);
  }
  double rew = terms[0] + terms[1] + terms[2] + terms[3] + terms[4] + terms[5] + terms[6];

  // :250  observation
  float* obs = obs_row;
  obs[0] = (float)bx;                                   // _get_observation (:651-670), sharing R and the heading
  obs[1] = (float)by;
  obs[2] = (float)b.v[0];
  obs[3] = (float)b.v[1];
  obs[4] = (float)b.w[2];
  obs[5] = (float)heading;
  for (int k = 0; k < p.num_obstacles; k++) {
    obs[6 + 2 * k] = (float)((double)c.f(SALP_F_OBSTACLE0_X + 2 * k) - b.pw[0]);
    obs[7 + 2 * k] = (float)((double)c.f(SALP_F_OBSTACLE0_X + 2 * k + 1) - b.pw[1]);
  }
  if (p.randomization & SALP_RAND_OBSERVATION) {     // _randomize_observations (salp_robot_env.py:183-194)
    uint32_t r[8];
    rand_block(cx.rc.seed, cx.rc.gid, cx.rc.episode, cx.rc.cycle, SALP_RNG_OBS, r);
    rand_block(cx.rc.seed, cx.rc.gid, cx.rc.episode, cx.rc.cycle, SALP_RNG_OBS + 1, r + 4);
    const float unc[6] = {0.05f, 0.05f, 0.2f, 0.2f, 0.02f, 0.1f};
    for (int k = 0; k < 6; k++) obs[k] = randomize_scalar_default_bounds(obs[k], unc[k], r[k]);
  }

  // :258-276  termination (cumulative, not exclusive)
  bool done = false, trunc = false;
  if (dist < p.target_radius) { done = true; rew += p.success_bonus; }
  else if (dist > p.out_of_bounds_distance) { trunc = true; rew -= p.out_of_bounds_penalty; }
  if (hit) { trunc = true; rew -= p.collision_penalty; }
  if (cycle >= p.max_cycles) { trunc = true; rew -= p.timeout_penalty; }
  if (crashed) {
    done = false;
    trunc = true;
    for (int k = 0; k < 7; k++) terms[k] = 0.0;
    rew = -p.out_of_bounds_penalty;
    for (int k = 0; k < D; k++) obs[k] = isfinite(obs[k]) ? obs[k] : 0.0f;
  }

  // episode bookkeeping (:199, :247-248, Monitor)
  c.n(SALP_F_EP_LENGTH) = ep_len + 1;
  c.d(SALP_F_EP_SUM_A0) += (double)a0;
  c.d(SALP_F_EP_SUM_A1) += (double)a1;
  c.d(SALP_F_EP_SUM_ABS_A2) += fabs((double)a2);
  for (int k = 0; k < 7; k++) c.d(SALP_F_EP_SUM_TERM0 + k) += terms[k];
  c.d(SALP_F_EP_RETURN) += rew;
  c.d(SALP_F_EP_SUBSTEPS) += (double)(K < 0 ? SALP_MAX_SUBSTEPS : K);
  c.f(SALP_F_PREV_ACTION0) = a0;
  c.f(SALP_F_PREV_ACTION1) = a1;
  c.f(SALP_F_PREV_ACTION2) = a2;

  const bool ended = done || trunc;
  if (io.episode_metrics && ended) {
    double* m = io.episode_metrics + i * SALP_NUM_EPISODE_METRICS;
    write_episode_metrics(c, b.pw[0], b.pw[1], dist, ep_len + 1, m);
    if (crashed) {
      for (int k = 0; k < SALP_NUM_EPISODE_METRICS; k++) m[k] = isfinite(m[k]) ? m[k] : 0.0;
      m[SALP_EM_NONFINITE] = 1.0;
    }
  }
  io.reward[i] = (float)rew;
  io.terminated[i] = done ? 1 : 0;
  io.truncated[i] = trunc ? 1 : 0;
  if (io.reward_terms) {
    for (int k = 0; k < 7; k++) io.reward_terms[8 * i + k] = terms[k];
    io.reward_terms[8 * i + 7] = rew;
  }
  if (io.substeps) io.substeps[i] = K < 0 ? SALP_MAX_SUBSTEPS : K;
  if (tobs_row)
    for (int k = 0; k < D; k++) tobs_row[k] = obs[k];
  // SB3 VecEnv worker semantics: reset the finished env, hand back the post-reset observation
  if ((flags & SALP_STEP_AUTORESET) && ended) env_reset(p, v, i, obs);
}

// History feed (salp_trace_cycle): the coming cycle of env i in float64 reference arithmetic on a
// scratch copy of its state; one row per substep, taken after Robot.step() (robot.py:757-766).
SALP_HD int env_trace_cycle(const SalpParams& p, const SalpView& v, int64_t i, float a0, float a1, float a2,
                            double* trace, int capacity) {
  Cols c{v, i};
  CyclePlan plan = make_cycle_plan(p, a0, a1, a2, c.d(SALP_F_NOZZLE_ANGLE1), c.d(SALP_F_NOZZLE_ANGLE2));
  Body64 b;
  load_body(c, b);
  refresh_shape_f64(p, b);
  b.mass_rate = (b.water_mass - b.prev_volume * p.density) / p.dt;
  double t = 0.0;
  int K = 0;
  while (cycle_running(plan, t) && K < SALP_MAX_SUBSTEPS) {
    substep_f64(p, plan, b, t);
    if (K < capacity) {
      double* r = trace + (int64_t)SALP_TRACE_WIDTH * K;
      for (int k = 0; k < 3; k++) { r[k] = b.pw[k]; r[3 + k] = b.eul[k]; r[6 + k] = b.v[k]; r[9 + k] = b.w[k]; }
      r[12] = b.length;
      r[13] = b.width;
    }
    K++;
  }
  return K;
}

template <int PREC>
SALP_HD void env_step(const SalpParams& p, const SalpDerived& dv, const SalpView& v, const SalpStepIO& io,
                      uint32_t flags, int64_t i, float* obs_row = nullptr, float* tobs_row = nullptr) {
  StepCtx cx;
  Body64 b;
  env_step_begin(p, v, io, i, cx, b);
  const double pos0[3] = {b.pos[0], b.pos[1], b.pos[2]};
  const double ang0[3] = {b.ang[0], b.ang[1], b.ang[2]};
  double t = 0.0;
  const int K = run_cycle<PREC>(p, dv, cx.plan, v.time_table, b, t, &cx.rc);
  env_step_end(p, v, io, flags, i, cx, pos0, ang0, b, K, t, obs_row, tobs_row);
}
