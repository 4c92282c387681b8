"""Shared parity machinery: replay a golden trace (or drive two backends side by side).

A *backend* is anything with the small duck-typed surface below; both the CPU oracle
(oracle.salp_oracle.OracleVecEnv) and the CUDA product (grasp_lab_salp_b200.SalpBatch)
provide it, so every parity test reads the same for both:

    .num_envs  .obs_dim
    .set_scene_pool(targets[N,P,2], obstacles[N,P,n,2])
    .reset(mask=None) -> obs[N,D]
    .step(actions[N,3], auto_reset=False) -> (obs, reward, terminated, truncated)   (numpy, host)
    .terms [N,8]  .substeps [N]  .metrics [N,20]
    .get_state(name) -> column[N]
"""
from __future__ import annotations

import os

import numpy as np

from conftest import GOLDEN_DIR

# golden "state" column name -> backend state column (None = not a product state column)
STATE_MAP = {
    "posw_x": "posw_x", "posw_y": "posw_y", "posw_z": "posw_z",
    "vel_x": "vel_x", "vel_y": "vel_y", "vel_z": "vel_z",
    "euler_x": "euler_x", "euler_y": "euler_y", "euler_z": "euler_z",
    "angvel_x": "angvel_x", "angvel_y": "angvel_y", "angvel_z": "angvel_z",
    "length": "length", "width": "width",
    "nozzle_angle1": "nozzle_angle1", "nozzle_angle2": "nozzle_angle2", "prev_dist": "prev_dist",
    "pos_x": "pos_x", "pos_y": "pos_y", "pos_z": "pos_z",
    "angle_x": "angle_x", "angle_y": "angle_y", "angle_z": "angle_z",
    "acc_x": "acc_x", "acc_y": "acc_y", "acc_z": "acc_z",
    "angacc_x": "angacc_x", "angacc_y": "angacc_y", "angacc_z": "angacc_z", "com_x": "com_x",
}
# near-zero channels (out-of-plane motion, SURVEY.md hard part 5): compared with an absolute floor
SMALL_CHANNELS = {"posw_z", "vel_z", "euler_x", "euler_y", "angvel_x", "angvel_y", "pos_z", "angle_x",
                  "angle_y", "acc_z", "angacc_x", "angacc_y"}
ANGLE_CHANNELS = {"euler_z", "angle_z"}
METRIC_COLS = {"path_length": 2, "direct_distance": 3, "path_efficiency": 4, "final_distance": 5,
               "initial_distance": 6, "avg_compression": 7, "avg_coast_time": 8, "avg_nozzle_angle": 9,
               "avg_velocity": 10, "avg_rewards_track": 11, "avg_rewards_heading": 12,
               "avg_rewards_smooth": 13, "avg_rewards_yaw": 14, "avg_rewards_time": 15,
               "avg_rewards_sideslip": 16, "avg_rewards_obstacle": 17}


def load_golden(name):
    """Golden trace as a dict-like.  The long-horizon sets (tools/gen_golden_long.py) store the
    post-reset observations and the episode metrics sparsely (one row per ended episode) and have
    no per-component reward terms; they are densified here to the layout replay_golden reads."""
    g = np.load(os.path.join(GOLDEN_DIR, name), allow_pickle=False)
    if "reset_env" not in g.files:
        return g
    g = {k: g[k] for k in g.files}
    e, t = g.pop("reset_env"), g.pop("reset_t")
    reset_obs = g["obs"].copy()
    reset_obs[e, t] = g.pop("reset_obs_rows")
    g["reset_obs"] = reset_obs
    rows = g.pop("metrics_rows")
    if len(rows) == len(e):
        n, T = g["K"].shape
        metrics = np.full((n, T, rows.shape[1]), np.nan)
        metrics[e, t] = rows
        g["metrics"] = metrics
    return g


def golden_params(g, **kw):
    from grasp_lab_salp_b200.params import default_params
    return default_params(refill_poly=g["refill_poly"], jet_poly=g["jet_poly"], **kw)


def rel_err(a, b, floor):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), floor)


# Tolerance sets.  rel_err(a, b, floor) = |a - b| / max(|b|, floor).
#  * F64 ("reference mode", and the C oracle): same float64 algorithm, different libm / summation
#    order -> 1e-9 relative.
#  * MIXED (fp32 motion state): the north-star tolerance, 1e-5 relative PER ENV-STEP on position,
#    velocity, heading, body shape, observation and reward, measured from identical state
#    (lockstep_compare(resync=True)).  SURVEY 8d asks for "relative 1e-5 with an absolute floor of 1e-7
#    for the near-zero z / roll / pitch channels".  A pure relative test is not meaningful for a
#    coordinate that happens to pass through zero (x, y, yaw do, all the time), and 1e-12 absolute on
#    roll is below what fp32 can carry (roll ~ 1e-3 rad, fp32 eps 6e-8).  So every channel gets the
#    SMALLEST floor with which the fp32 kernel holds 1e-5 -- MIXED_FLOORS below, measured on 512 envs
#    x 24 steps of uniform and clipped-Gaussian actions (fraction of samples that pass the pure
#    relative test / largest absolute error / floor needed), DESIGN.md section 3.2:
#        x, y            99.1 %   1.2e-6 m      0.053 m   (one cycle moves the body 0.1-0.7 m)
#        vx, vy          98.7 %   3.0e-7 m/s    0.0092 m/s
#        yaw             99.8 %   1.3e-6 rad    0.095 rad
#        yaw rate       100   %   3.7e-6 rad/s  0
#        z               95.2 %   3.4e-9 m      3.8e-5 m  (|z| ~ 1e-3 m: driven by rounding-level asymmetries)
#        roll, pitch     95.0 %   5.1e-9 rad    3.5e-4 rad
#        vz              84.6 %   1.4e-10 m/s   4.1e-7 m/s
#        length, width, volume, centre of mass: identical bits (computed in fp64 in both)
#    i.e. the floors are 2-3x the "needed" column.  Rewards carry the reference's own x100 gain on
#    distances (salp_robot_env.py:352), hence floor 10.  Accelerations are not in the north-star list
#    and are sums of cancelling forces: 1e-4.
TOL_F64 = dict(rtol=1e-9, small_rtol=1e-6, small_floor=1e-4)
MIXED_FLOORS = {
    "posw_x": 0.1, "posw_y": 0.1, "pos_x": 0.1, "pos_y": 0.1,
    "vel_x": 0.02, "vel_y": 0.02,
    "euler_z": 0.2, "angle_z": 0.2, "angvel_z": 1e-3,
    "posw_z": 1e-4, "pos_z": 1e-4, "vel_z": 2e-6, "euler_x": 1e-3, "euler_y": 1e-3, "angle_x": 1e-3, "angle_y": 1e-3,
    "angvel_x": 1e-7, "angvel_y": 1e-7,
    "length": 1e-9, "width": 1e-9, "prev_volume": 1e-9, "com_x": 1e-9, "prev_dist": 0.1,
}
# Golden traces are FREE-RUNNING for up to 30 env-steps (no re-synchronisation), so fp32 rounding
# accumulates as a random walk on the neutrally stable channels (position, heading): 3e-5 there;
# the per-step figure of 1e-5 is tested from identical state by lockstep_compare() below.
TOL_MIXED_FREE_RUN = dict(rtol=3e-5, floor=0.1, small_floor=0.1, accel_rtol=1e-4, accel_floor=1.0,
                          reward_floor=10.0, obs_floor=0.1, metric_floor=0.1)
TOL_MIXED = dict(rtol=1e-5, floor=0.1, small_floor=0.1, accel_rtol=1e-4, accel_floor=1.0, reward_floor=10.0,
                 obs_floor=0.1, metric_floor=0.1)


def replay_golden(backend, g, rtol, *, small_rtol=None, small_floor=1e-7, floor=1e-9, accel_rtol=None,
                  accel_floor=1e-6, reward_floor=1e-3, obs_floor=1e-4, metric_floor=1e-6,
                  check_metrics=True, report=None, regular_tilt=None):
    """Drive `backend` with the golden actions/scenes and compare every recorded quantity.

    Integer/flag quantities (K, cycle, phase, terminated, truncated, hence reset indices)
    must be bit-exact; floats within `rtol` relative (absolute floor `floor`; `small_floor`
    for the near-zero out-of-plane channels).  Returns the worst relative error seen.

    regular_tilt: the model's out-of-plane motion is unstable -- |roll|, |pitch| grow about an
    e-fold per 25 cycles in the reference -- and once they reach O(1) rad the Euler-rate matrix
    (dynamics.py:21-31) passes its pitch = +-pi/2 singularity and the trajectory is chaotic: the
    reference does not reproduce ITSELF there under a 1e-15 perturbation.  With
    regular_tilt set, floats of an env are compared only while the GOLDEN max(|roll|, |pitch|) of
    its current episode has stayed below that angle; integers and flags are compared always.
    """
    actions = g["actions"]
    n, T, _ = actions.shape
    assert backend.num_envs == n
    names = [str(s) for s in g["state_names"]]
    backend.set_scene_pool(g["targets"], g["obstacles"])
    obs0 = backend.reset()
    np.testing.assert_allclose(obs0, g["first_obs"], rtol=max(rtol, 1e-6), atol=1e-6)
    worst = {}
    regular = np.ones(n, bool)
    compared = chaotic_flag_mismatch = 0

    def track(key, err):
        worst[key] = max(worst.get(key, 0.0), float(np.max(err)) if np.size(err) else 0.0)

    for t in range(T):
        obs, rew, term, trunc = backend.step(actions[:, t], auto_reset=False)
        ctx = f"step {t}"
        if regular_tilt is not None:
            tilt = np.maximum(np.abs(g["state"][:, t, names.index("euler_x")]),
                              np.abs(g["state"][:, t, names.index("euler_y")]))
            regular &= tilt < regular_tilt
        compared += int(regular.sum())
        np.testing.assert_array_equal(backend.substeps, g["K"][:, t], err_msg=f"K {ctx}")
        np.testing.assert_array_equal(backend.get_state("cycle"), g["cycle"][:, t], err_msg=f"cycle {ctx}")
        np.testing.assert_array_equal(backend.get_state("phase"), g["phase"][:, t], err_msg=f"phase {ctx}")
        g_term, g_trunc = g["terminated"][:, t].astype(bool), g["truncated"][:, t].astype(bool)
        # (in the chaotic regime the position is not reproducible, so neither are the position-dependent
        #  flags; the cycle-count truncation of salp_robot_env.py:274-276 is, and is checked everywhere)
        np.testing.assert_array_equal(term.astype(bool)[regular], g_term[regular], err_msg=f"terminated {ctx}")
        np.testing.assert_array_equal(trunc.astype(bool)[regular], g_trunc[regular], err_msg=f"truncated {ctx}")
        timed_out = g["cycle"][:, t] >= backend.params.max_cycles
        assert trunc.astype(bool)[timed_out].all(), f"cycle >= max_cycles must truncate, {ctx}"
        chaotic_flag_mismatch += int(((term.astype(bool) != g_term) | (trunc.astype(bool) != g_trunc))[~regular].sum())
        for j, nm in enumerate(names):
            col = STATE_MAP.get(nm)
            if col is None and nm != "volume":
                continue
            ref = g["state"][:, t, j]
            if nm == "volume":      # Robot.volume (robot.py:1055-1056) is not carried: ellipsoid(length, width) - tube
                lh, wh = 0.5 * backend.get_state("length"), 0.5 * backend.get_state("width")
                got = (4.0 / 3.0) * np.pi * lh * wh * wh - backend.params.tube_volume
            else:
                got = backend.get_state(col)
            fl = small_floor if nm in SMALL_CHANNELS else floor
            if nm in ANGLE_CHANNELS and floor >= 0.1:
                fl = max(fl, 1.0)
            tol = (small_rtol or rtol) if nm in SMALL_CHANNELS else rtol
            if nm.startswith(("acc_", "angacc_")):
                # accelerations are finite-difference driven (differences of O(1) numbers / dt):
                # the absolute floor scales with 1/dt^2 of the geometry's rounding noise
                fl = max(fl, accel_floor)
                tol = max(tol, accel_rtol or 0.0)
            finite = np.isfinite(ref)
            np.testing.assert_array_equal(np.isfinite(got), finite, err_msg=f"{nm} finiteness {ctx}")
            finite = finite & regular
            e = rel_err(got[finite], ref[finite], fl)
            track(nm, e)
            assert np.all(e <= tol), f"{nm} {ctx}: rel err {e.max():.3e} > {tol:g} (got {got}, ref {ref})"
        fin = np.isfinite(g["reward"][:, t]) & regular
        e = rel_err(rew[fin], g["reward"][:, t][fin], reward_floor)
        track("reward", e)
        assert np.all(e <= max(rtol, 1e-6) * 10), f"reward {ctx}: {e.max():.3e}"
        if "terms" in g:
            tf = np.isfinite(g["terms"][:, t]) & regular[:, None]
            e = rel_err(backend.terms[:, :7][tf], g["terms"][:, t][tf], reward_floor)
            track("reward_terms", e)
            assert np.all(e <= max(rtol, 1e-6) * 10), f"reward terms {ctx}: {e.max():.3e}"
        of = np.isfinite(g["obs"][:, t]) & regular[:, None]
        e = rel_err(obs[of], g["obs"][:, t][of], obs_floor)
        track("obs", e)
        assert np.all(e <= max(rtol, 2e-7) * 10), f"obs {ctx}: {e.max():.3e}"
        ended = g_term | g_trunc          # resets follow the golden episode structure (== the backend's wherever flags are compared)
        if check_metrics and "metrics" in g and ended.any():
            keys = [str(k) for k in g["metric_keys"]]
            for j, k in enumerate(keys):
                ref = g["metrics"][ended, t, j]
                got = backend.metrics[ended, METRIC_COLS[k]]
                ok = np.isfinite(ref) & regular[ended]
                e = rel_err(got[ok], ref[ok], metric_floor)
                track("metrics", e)
                assert np.all(e <= max(rtol, 1e-6) * 10), f"metric {k} {ctx}: {e.max():.3e}"
        if ended.any():
            obs_r = backend.reset(mask=ended.astype(np.uint8))
            np.testing.assert_allclose(obs_r[ended], g["reset_obs"][ended, t], rtol=1e-6, atol=1e-7,
                                       err_msg=f"post-reset obs {ctx}")
            regular[ended] = True
    worst["float_rows_compared"] = compared
    worst["chaotic_flag_mismatches"] = chaotic_flag_mismatch
    if report is not None:
        report.update(worst)
    return max(v for k, v in worst.items() if k not in ("float_rows_compared", "chaotic_flag_mismatches"))


# ---------------------------------------------------------------------------------------------
# lockstep comparison of two backends (product vs oracle) on the same actions
# ---------------------------------------------------------------------------------------------
ALL_COLUMNS = None


def all_columns():
    global ALL_COLUMNS
    if ALL_COLUMNS is None:
        from grasp_lab_salp_b200.params import FIELDS
        ALL_COLUMNS = [n for n in FIELDS if n != "speed_world" and not n.startswith("ou_")] + ["speed_world"]
    return ALL_COLUMNS


PER_STEP_CHANNELS = ["posw_x", "posw_y", "vel_x", "vel_y", "euler_z", "angvel_z", "length", "width",
                     "posw_z", "vel_z", "euler_x", "euler_y", "angvel_x", "angvel_y", "pos_x", "pos_y", "angle_z",
                     "prev_volume", "com_x", "prev_dist"]


def obs_errors(got, ref, floor, tag):
    """Observation error metrics (salp_robot_env.py:651-670 layout).  The body-frame target vector
    obs[0:2] and each obstacle vector are compared as VECTORS (|d| / max(|ref vec|, floor)): a yaw
    error rotates the vector, which is a large relative error on whichever component happens to
    be small.  obs[5] = atan2(d_body.y, d_body.x) amplifies a position error by 1/distance
    (terminal observations are taken inside the 0.2 m target radius), so it is compared as an arc
    length |d heading| * min(distance, 1) against a 1 rad scale.  obs[2:5] (v_x, v_y, w_z) directly."""
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    ok = np.isfinite(ref).all(axis=1)
    got, ref = got[ok], ref[ok]
    out = {}
    if not len(ref):
        return {tag: 0.0}
    worst = 0.0
    for a in [0] + list(range(6, ref.shape[1], 2)):
        d = np.hypot(got[:, a] - ref[:, a], got[:, a + 1] - ref[:, a + 1])
        nrm = np.hypot(ref[:, a], ref[:, a + 1])
        worst = max(worst, float((d / np.maximum(nrm, floor)).max()))
    worst = max(worst, float(rel_err(got[:, 2:5], ref[:, 2:5], floor).max()))
    out[tag] = worst
    dist = np.hypot(ref[:, 0], ref[:, 1])
    dh = np.abs((got[:, 5] - ref[:, 5] + np.pi) % (2 * np.pi) - np.pi)
    out[tag + "_heading_arc"] = float((dh * np.minimum(dist, 1.0)).max())
    return out


def lockstep_compare(product, oracle, actions, *, resync, rtol, floor, reward_floor=10.0, obs_floor=0.1,
                     small_floor=None, num_obstacles=2, report=None, floors=None):
    """Step `product` and `oracle` with the same actions [T,N,3] (auto-reset on, same scenes).

    Bit-exact: K, cycle, phase, terminated, truncated (hence the reset indices), episode index.
    Floats: |a-b| <= rtol * max(|b|, floor) on PER_STEP_CHANNELS, obs, reward.
    resync=True copies the oracle's full state into the product before every step, so each
    comparison is a single env-step from identical state (the per-step tolerance);
    resync=False lets both run free (long-horizon drift; returns the error history instead of
    asserting on floats).
    """
    T = actions.shape[0]
    obs_p = product.reset().copy()
    obs_o = oracle.reset().copy()
    np.testing.assert_allclose(obs_p, obs_o, rtol=1e-6, atol=1e-6)
    worst = {}
    history = []
    obstacle_cols = [f"obstacle{i}_{a}" for i in range(num_obstacles) for a in "xy"]
    skip = {f"obstacle{i}_{a}" for i in range(num_obstacles, 8) for a in "xy"}
    for t in range(T):
        if resync and t > 0:
            for col in all_columns():
                if col in skip:
                    continue
                product.set_state(col, oracle.get_state(col))
        a = actions[t]
        obs_p, rew_p, term_p, trunc_p = product.step(a, auto_reset=True)
        obs_o, rew_o, term_o, trunc_o = oracle.step(a, auto_reset=True)
        ctx = f"step {t}"
        np.testing.assert_array_equal(product.substeps, oracle.substeps, err_msg=f"K {ctx}")
        np.testing.assert_array_equal(term_p.astype(np.uint8), term_o.astype(np.uint8), err_msg=f"terminated {ctx}")
        np.testing.assert_array_equal(trunc_p.astype(np.uint8), trunc_o.astype(np.uint8), err_msg=f"truncated {ctx}")
        for col in ("cycle", "phase", "ep_length", "episode_index"):
            np.testing.assert_array_equal(product.get_state(col), oracle.get_state(col), err_msg=f"{col} {ctx}")
        for col in ["target_x", "target_y"] + obstacle_cols:
            np.testing.assert_array_equal(product.get_state(col), oracle.get_state(col), err_msg=f"{col} {ctx}")
        errs = {}
        for col in PER_STEP_CHANNELS:
            ref = oracle.get_state(col)
            got = product.get_state(col)
            ok = np.isfinite(ref)
            # out-of-plane channels are excited only by rounding-level asymmetries (nozzle direction
            # z ~ 2e-16): they are noise-driven and compared against an absolute floor
            if floors is not None:
                fl = floors[col]        # per-channel floors (MIXED_FLOORS: measured, see the table above)
            else:
                fl = max(floor, small_floor or 1e-4) if col in SMALL_CHANNELS else floor
                if col in ANGLE_CHANNELS and floor >= 0.1:
                    fl = max(fl, 1.0)   # an angle is relative to 1 rad, not to how close to 0 it happens to end
            e = rel_err(got[ok], ref[ok], fl)
            errs[col] = float(e.max()) if e.size else 0.0
        rew_o64, rew_p64 = oracle.terms[:, 7], product.terms[:, 7]      # float64 totals (io.reward is float32)
        ok = np.isfinite(rew_o64)
        errs["reward"] = float(rel_err(rew_p64[ok], rew_o64[ok], reward_floor).max()) if ok.any() else 0.0
        np.testing.assert_allclose(rew_p[ok], rew_o64[ok].astype(np.float32), rtol=1e-5, atol=rtol * reward_floor * 10)
        errs.update(obs_errors(product.terminal_obs, oracle.terminal_obs, obs_floor, "obs"))
        errs.update(obs_errors(obs_p, obs_o, obs_floor, "obs_after_reset"))
        history.append(errs)
        for k, v in errs.items():
            worst[k] = max(worst.get(k, 0.0), v)
        if resync:
            bad = {k: v for k, v in errs.items() if not v <= rtol}
            assert not bad, f"{ctx}: per-step error above {rtol:g}: {bad}"
    if report is not None:
        report.update(worst)
    return history


def sample_scene_pool(rng, n_envs, P, num_obstacles=2, lo=(-2.0, -1.5), hi=(2.0, 1.5)):
    """Host-side scene sampler with the reference's rejection rule (salp_robot_env.py:535-559)."""
    targets = np.zeros((n_envs, P, 2), np.float32)
    obstacles = np.zeros((n_envs, P, num_obstacles, 2), np.float32)
    for i in range(n_envs):
        for s in range(P):
            t = rng.uniform(lo, hi).astype(np.float32)
            obs = []
            while len(obs) < num_obstacles:
                pos = rng.uniform(lo, hi).astype(np.float32)
                if (np.linalg.norm(pos) > 0.5 and np.linalg.norm(pos - t) > 0.5
                        and not any(np.linalg.norm(pos - o) < 0.5 for o in obs)):
                    obs.append(pos)
            targets[i, s] = t
            obstacles[i, s] = np.array(obs).reshape(num_obstacles, 2)
    return targets, obstacles


def check_blowup_golden(make_backend):
    """tests/golden/ref_blowup.npz: 512 single cycles from rest (carried nozzle angles injected)
    with contractions around the zero crossings of the reference's refill/jet-time polynomials;
    in 16 of them the live reference raised LinAlgError (its state overflowed).  The backend
    must cut exactly those episodes (truncated, reward -200, metric `nonfinite` = 1, finite
    observation) and reproduce K and the final pose of all the others."""
    g = load_golden("ref_blowup.npz")
    n = len(g["actions"])
    env = make_backend(n, g)
    env.set_scene_pool(np.tile(np.array([[[1.8, 1.2]]], np.float32), (n, 1, 1)),
                       np.tile(np.array([[[[-1.5, -1.0], [1.5, -1.0]]]], np.float32), (n, 1, 1, 1)))
    env.reset()
    env.set_state("nozzle_angle1", g["angle1"])
    env.set_state("nozzle_angle2", g["angle2"])
    obs, rew, term, trunc = env.step(g["actions"])
    raised = g["raised"].astype(bool)
    cut = (env.metrics[:, 19] == 1.0) & trunc.astype(bool)
    np.testing.assert_array_equal(cut, raised)
    assert np.isfinite(obs).all() and np.isfinite(rew).all()
    np.testing.assert_array_equal(term[raised], 0)
    np.testing.assert_allclose(rew[raised], -200.0)
    ok = ~raised
    np.testing.assert_array_equal(env.substeps[ok], g["substeps_run"][ok])
    got = np.stack([env.get_state("posw_x"), env.get_state("posw_y"), env.get_state("euler_z")], axis=1)
    return float(rel_err(got[ok], g["final"][ok], 0.1).max())
