#!/usr/bin/env python
"""Generate the LONG-HORIZON golden sets from the LIVE, UNMODIFIED Python reference
(round-2 parity pins; same harness and SB3-worker driving as tools/gen_golden.py).

    python tools/gen_golden_long.py config2    # tests/golden/ref_config2.npz   256 envs x 200 steps, auto-reset
    python tools/gen_golden_long.py long       # tests/golden/ref_long.npz      24 episodes that reach cycle 500
    python tools/gen_golden_long.py config1    # tests/golden/ref_config1.npz   1 env, 1000 uniform-random steps
    python tools/gen_golden_long.py ksweep     # tests/golden/ref_ksweep.npz    1 048 576 actions: K and IK only

Needs /root/reference (or $SALP_REF_DIR / baseline/_ref), numpy, numba: build container only.
The files keep the layout tests/parity.py:replay_golden reads, with a reduced state-column set,
narrow integer dtypes and the post-reset observations stored sparsely (only where an episode ended).

  * config2 -- BASELINE config 2's Python-pinned subsample (>= 256 envs x >= 200 steps,
    uniform Box actions, injected scenes).
  * long    -- action sequences, found with the C oracle (circling / biased-yaw policies), whose first
    episode survives to the `robot.cycle >= 500` truncation (salp_robot_env.py:274-276) and tumbles on
    the way (|roll|, |pitch| > 0.3 rad): pins the timeout flag and the large-angle regime of
    robot.py:860-875 against Python.  The SEARCH uses oracle/ (fast); what is RECORDED is the reference.
  * config1 -- BASELINE config 1 (test_simple.py:22-39 construction, 1000 uniform-random steps).
  * ksweep  -- the reference's real `while self.cycle_time < total` loop (robot.py:756-757) with
    Robot.step replaced by its two time-advancing statements (robot.py:674-675), driven through the
    real SalpRobotEnv.step (float32 rescale, IK, set_control): K, angle1, angle2, turn_time for 2^20
    actions in 1024 chains of 1024 (the nozzle angles carry from action to action).
"""
from __future__ import annotations

import os
import sys
import time
from multiprocessing import Pool

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from oracle import ref_harness as rh  # noqa: E402
from gen_golden import METRIC_KEYS, sample_scenes  # noqa: E402

COMPACT_STATE = ["posw_x", "posw_y", "vel_x", "vel_y", "euler_z", "angvel_z", "volume"]
LONG_STATE = ["posw_x", "posw_y", "posw_z", "vel_x", "vel_y", "vel_z", "euler_x", "euler_y", "euler_z",
              "angvel_x", "angvel_y", "angvel_z", "length", "width", "volume", "nozzle_angle1", "nozzle_angle2"]
NPROC = int(os.environ.get("SALP_GEN_PROCS", "8"))


def _state(env, names):
    r = env.robot
    full = dict(posw_x=r.position_world[0], posw_y=r.position_world[1], posw_z=r.position_world[2],
                vel_x=r.velocity[0], vel_y=r.velocity[1], vel_z=r.velocity[2],
                euler_x=r.euler_angle[0], euler_y=r.euler_angle[1], euler_z=r.euler_angle[2],
                angvel_x=r.angular_velocity[0], angvel_y=r.angular_velocity[1], angvel_z=r.angular_velocity[2],
                length=r.length, width=r.width, volume=r.volume,
                nozzle_angle1=r.nozzle.angle1, nozzle_angle2=r.nozzle.angle2)
    return np.array([full[n] for n in names], np.float64)


def run_env_compact(args):
    """One reference env, T steps, SB3-worker auto-reset with injected scenes; compact record."""
    actions, targets, obstacles, names, want_metrics = args
    T, P = actions.shape[0], targets.shape[0]
    env = rh.make_env()
    counter = {"k": 0}
    orig_step = env.robot.step

    def counting_step():
        counter["k"] += 1
        orig_step()

    env.robot.step = counting_step
    env.reset()
    episode = 0
    obs0 = rh.inject_scene(env, targets[0], obstacles[0])
    D = obs0.shape[0]
    out = dict(obs=np.zeros((T, D), np.float32), reward=np.zeros(T), terminated=np.zeros(T, np.uint8),
               truncated=np.zeros(T, np.uint8), K=np.zeros(T, np.int16), cycle=np.zeros(T, np.int16),
               phase=np.zeros(T, np.int8), state=np.zeros((T, len(names))), first_obs=obs0)
    reset_t, reset_obs, metrics = [], [], []
    for t in range(T):
        counter["k"] = 0
        obs, rew, done, trunc, info = env.step(actions[t].copy())
        r = env.robot
        out["obs"][t] = obs
        out["reward"][t] = rew
        out["terminated"][t] = done
        out["truncated"][t] = trunc
        out["K"][t] = counter["k"]
        out["cycle"][t] = r.cycle
        out["phase"][t] = r.state.value
        out["state"][t] = _state(env, names)
        if done or trunc:
            if want_metrics:
                metrics.append([info.get(k, np.nan) for k in METRIC_KEYS])
            episode += 1
            env.reset()
            reset_t.append(t)
            reset_obs.append(rh.inject_scene(env, targets[episode % P], obstacles[episode % P]))
    out["reset_t"] = np.array(reset_t, np.int32)
    out["reset_obs_rows"] = np.array(reset_obs, np.float32).reshape(len(reset_t), D)
    out["metrics_rows"] = np.array(metrics, np.float64).reshape(len(metrics), len(METRIC_KEYS))
    return out


def write_compact(name, actions, targets, obstacles, names, note, want_metrics=True):
    n = actions.shape[0]
    t0 = time.time()
    with Pool(min(NPROC, n)) as pool:
        res = pool.map(run_env_compact, [(actions[i], targets[i], obstacles[i], names, want_metrics)
                                         for i in range(n)], chunksize=1)
    dense = {k: np.stack([r[k] for r in res]) for k in res[0] if k not in ("reset_t", "reset_obs_rows", "metrics_rows")}
    reset_env = np.concatenate([np.full(len(r["reset_t"]), i, np.int32) for i, r in enumerate(res)])
    reset_t = np.concatenate([r["reset_t"] for r in res])
    m = rh.load()
    path = os.path.join(ROOT, "tests", "golden", name)
    np.savez_compressed(
        path, actions=actions, targets=targets, obstacles=obstacles,
        refill_poly=m.geometry.fit_compression_refill_time_relation_jit(),
        jet_poly=m.geometry.fit_compression_propulsion_time_relation_jit(),
        state_names=np.array(names), metric_keys=np.array(METRIC_KEYS), note=np.array(note),
        reset_env=reset_env, reset_t=reset_t,
        reset_obs_rows=np.concatenate([r["reset_obs_rows"] for r in res]),
        metrics_rows=np.concatenate([r["metrics_rows"] for r in res]), **dense)
    ended = dense["terminated"] | dense["truncated"]
    print(f"{name}: {n} envs x {actions.shape[1]} steps in {time.time() - t0:.0f} s, mean K {dense['K'].mean():.1f}, "
          f"episodes ended {int(ended.sum())}, max cycle {int(dense['cycle'].max())}, "
          f"size {os.path.getsize(path) / 1e6:.2f} MB", flush=True)


# ---------------------------------------------------------------------------------------------
# long episodes: candidate action sequences, screened with the C oracle
# ---------------------------------------------------------------------------------------------
def long_candidates(rng, kind, T):
    a = np.zeros((T, 3), np.float32)
    if kind == "circle":          # strong strokes, short coasts, nozzle held hard to one side
        s = rng.choice([-1, 1])
        a[:, 0] = rng.uniform(0.3, 1, T)
        a[:, 1] = rng.uniform(0, 0.1, T)
        a[:, 2] = np.clip(s * rng.uniform(0.6, 1.0) + rng.normal(0, 0.1, T), -1, 1)
    else:                         # "mix": uniform strokes, mostly short coasts with rare long ones, biased yaw
        s = rng.choice([-1, 1])
        a[:, 0] = rng.uniform(0, 1, T)
        a[:, 1] = rng.uniform(0, 1, T) ** 3
        a[:, 2] = np.clip(s * 0.8 + rng.normal(0, 0.3, T), -1, 1)
    return a


def find_long_episodes(n_circle=16, n_mix=8, T=520):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle.salp_oracle import OracleVecEnv
    from grasp_lab_salp_b200.params import default_params
    m = rh.load()
    params = default_params(refill_poly=m.geometry.fit_compression_refill_time_relation_jit(),
                            jet_poly=m.geometry.fit_compression_propulsion_time_relation_jit())
    rng = np.random.default_rng(20261019)
    target = np.array([2.0, 1.5], np.float32)
    obst = np.array([[-2.0, -1.5], [2.0, -1.5]], np.float32)
    picked = []
    for kind, want in (("circle", n_circle), ("mix", n_mix)):
        n = 96
        acts = np.stack([long_candidates(rng, kind, T) for _ in range(n)])
        env = OracleVecEnv(n, params, threads=8)
        env.set_scene_pool(np.tile(target[None, None], (n, 1, 1)), np.tile(obst[None, None], (n, 1, 1, 1)))
        env.reset()
        first_end = np.full(n, -1)
        tilt = np.zeros(n)
        for s in range(T):
            _, _, te, tr = env.step(acts[:, s], auto_reset=True)
            ended = (te | tr).astype(bool)
            alive = first_end < 0
            now = np.maximum(np.abs(env.get_state("euler_x")), np.abs(env.get_state("euler_y")))
            tilt = np.where(alive & ~ended, np.maximum(tilt, now), tilt)
            first_end = np.where(alive & ended, s + 1, first_end)
        ok = np.flatnonzero((first_end == 500) & (tilt > 0.3))
        print(f"long/{kind}: {len(ok)} of {n} candidates reach cycle 500 and tumble; taking {want}")
        assert len(ok) >= want
        picked += [acts[i] for i in ok[:want]]
        env.close()
    n = len(picked)
    return (np.stack(picked), np.tile(target[None, None], (n, 1, 1)), np.tile(obst[None, None], (n, 1, 1, 1)))


def screen_no_blowup(actions, targets, obstacles):
    """Run the C oracle over the planned trace; an action after which the oracle's state is non-finite
    or near the blow-up (|v| > 50 m/s) gets its contraction moved by +0.02 and the screen restarts.
    Returns the number of repaired actions (the trace in `actions` is edited in place)."""
    from oracle.salp_oracle import OracleVecEnv
    from grasp_lab_salp_b200.params import default_params
    m = rh.load()
    params = default_params(refill_poly=m.geometry.fit_compression_refill_time_relation_jit(),
                            jet_poly=m.geometry.fit_compression_propulsion_time_relation_jit())
    n, T = actions.shape[:2]
    repaired = 0
    while True:
        env = OracleVecEnv(n, params, threads=8)
        env.set_scene_pool(targets, obstacles)
        env.reset()
        bad = None
        for t in range(T):
            _, _, te, tr = env.step(actions[:, t], auto_reset=False)
            ended = (te | tr).astype(bool)
            speed = np.abs(np.stack([env.get_state("vel_x"), env.get_state("vel_y"), env.get_state("angvel_z")]))
            wild = ~np.isfinite(speed).all(axis=0) | (speed.max(axis=0) > 50.0) | (ended & (env.metrics[:, 19] == 1.0))
            if wild.any():
                bad = (np.flatnonzero(wild), t)
                break
            if ended.any():
                env.reset(mask=ended.astype(np.uint8))
        env.close()
        if bad is None:
            break
        for i in bad[0]:
            print(f"  repair: env {i} step {bad[1]} a0 = {actions[i, bad[1], 0]:.5f}", flush=True)
            actions[i, bad[1], 0] = np.float32(min(1.0, actions[i, bad[1], 0] + 0.02))
            repaired += 1
    print(f"screened {n} x {T} steps with the C oracle: no blow-up ({repaired} actions repaired)", flush=True)
    return repaired


# ---------------------------------------------------------------------------------------------
# K / IK sweep
# ---------------------------------------------------------------------------------------------
def hash_uniform(idx):
    """Counter-based uniform [0,1) float32 with 24 random bits (splitmix64 finaliser); depends on nothing
    but integer arithmetic, so the test regenerates the very same actions without storing them."""
    z = (np.asarray(idx, np.uint64) + np.uint64(0x9E3779B97F4A7C15)) * np.uint64(0xBF58476D1CE4E5B9)
    z ^= z >> np.uint64(30)
    z *= np.uint64(0x94D049BB133111EB)
    z ^= z >> np.uint64(27)
    z *= np.uint64(0xBF58476D1CE4E5B9)
    z ^= z >> np.uint64(31)
    return ((z >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)


def ksweep_actions(n_chains, T):
    """[n_chains, T, 3] float32.  Chains 0 mod 4, 1 mod 4, 2 mod 4: uniform on the Box (input A).
    Chains 3 mod 4: clipped -- a third of the entries pinned to each bound (input B-like).
    Every 64th chain replaces a0 by a value next to the zero crossings of the refill / jet time
    polynomials (a0 in [0.08, 0.1]) and every 64th+1 chain uses tiny yaws (|a2| < 1e-3)."""
    with np.errstate(over="ignore"):
        idx = np.arange(n_chains * T * 3, dtype=np.uint64).reshape(n_chains, T, 3)
        u = hash_uniform(idx)
    a = u.copy()
    a[:, :, 2] = u[:, :, 2] * np.float32(2) - np.float32(1)
    c = np.arange(n_chains)
    clip = (c % 4) == 3
    w = np.clip(u[clip] * np.float32(3) - np.float32(1), 0, 1).astype(np.float32)
    w[:, :, 2] = w[:, :, 2] * np.float32(2) - np.float32(1)
    a[clip] = w
    edge = (c % 64) == 0
    a[edge, :, 0] = np.float32(0.08) + u[edge, :, 0] * np.float32(0.02)
    tiny = (c % 64) == 1
    a[tiny, :, 2] = (u[tiny, :, 2] - np.float32(0.5)) * np.float32(2e-3)
    return a.astype(np.float32)


def run_chain(actions):
    env = rh.make_env()
    env.reset()
    rh.inject_scene(env, [2.0, 1.5], [[-2.0, -1.5], [2.0, -1.5]])
    r = env.robot
    counter = {"k": 0}

    def time_only_step():           # robot.py:674-675, nothing else of Robot.step
        counter["k"] += 1
        r.cycle_time += r.dt
        r.time += r.dt

    r.step = time_only_step
    T = actions.shape[0]
    K = np.zeros(T, np.int16)
    ang = np.zeros((T, 3))
    for t in range(T):
        counter["k"] = 0
        env.step(actions[t].copy())
        K[t] = counter["k"]
        ang[t] = (r.nozzle.angle1, r.nozzle.angle2, r.nozzle.turn_time)
        if t % 256 == 255:          # keep the env's per-episode python lists short
            env.episode_actions.clear(); env.episode_rewards.clear(); env.episode_reward_components.clear()
            env.episode_positions[:] = env.episode_positions[:1]
            env.episode_distances_to_target[:] = env.episode_distances_to_target[:1]
            env.episode_velocities[:] = env.episode_velocities[:1]
    return K, ang


def main():
    which = sys.argv[1:] or ["long", "config1", "ksweep", "config2"]
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)

    if "long" in which:
        a, t, o = find_long_episodes()
        write_compact("ref_long.npz", a, t, o, LONG_STATE,
                      "24 x 520 steps; first episode of every env survives to the cycle>=500 truncation and tumbles")

    if "config1" in which:
        rng = np.random.default_rng(1)
        a = rng.uniform([0, 0, -1], [1, 1, 1], size=(1, 1000, 3)).astype(np.float32)
        t, o = sample_scenes(rng, 1, 256)
        write_compact("ref_config1.npz", a, t, o, LONG_STATE, "BASELINE config 1: one env, 1000 uniform-random steps")

    if "ksweep" in which:
        n, T = 1024, 1024
        a = ksweep_actions(n, T)
        t0 = time.time()
        with Pool(NPROC) as pool:
            res = pool.map(run_chain, [a[i] for i in range(n)], chunksize=4)
        K = np.stack([r[0] for r in res])
        ang = np.stack([r[1] for r in res])
        sub = 64
        path = os.path.join(ROOT, "tests", "golden", "ref_ksweep.npz")
        m = rh.load()
        np.savez_compressed(
            path, K=K, n_chains=n, T=T, angles_first_chains=ang[:sub], angles_last_step=ang[:, -1],
            angle_checksum=np.array([np.abs(ang[:, :, 0]).sum(), np.abs(ang[:, :, 1]).sum(), ang[:, :, 2].sum()]),
            refill_poly=m.geometry.fit_compression_refill_time_relation_jit(),
            jet_poly=m.geometry.fit_compression_propulsion_time_relation_jit(),
            actions_probe=a[:2, :4],
            note=np.array("K [chain, step] of the reference's real cycle loop with Robot.step reduced to its time advance; "
                          "actions = tools/gen_golden_long.py:ksweep_actions(1024, 1024); angles (angle1, angle2, turn_time) "
                          f"for the first {sub} chains and for the last step of every chain"))
        print(f"ref_ksweep.npz: {n * T} actions in {time.time() - t0:.0f} s, mean K {K.mean():.1f}, K==0: {(K == 0).sum()}, "
              f"max K {K.max()}, size {os.path.getsize(path) / 1e6:.2f} MB", flush=True)

    if "config2" in which:
        n, T = 256, 200
        rng = np.random.default_rng(20261020)
        a = np.stack([rng.uniform([0, 0, -1], [1, 1, 1], size=(T, 3)) for _ in range(n)]).astype(np.float32)
        # Contractions whose jet_time is a fraction of one substep make the reference's integrator
        # diverge until np.linalg.solve raises LinAlgError (the process dies; that band has its own
        # golden set, ref_blowup.npz).  Here such draws (0.75 % of them) are moved out of the band so
        # that the reference survives all 200 steps; screened with the C oracle below.
        band = (a[..., 0] >= 0.088) & (a[..., 0] <= 0.0955)
        a[..., 0] = np.where(band, a[..., 0] + np.float32(0.01), a[..., 0])
        t, o = sample_scenes(rng, n, 64)
        screen_no_blowup(a, t, o)
        write_compact("ref_config2.npz", a, t, o, COMPACT_STATE,
                      "BASELINE config 2 Python pin: 256 envs x 200 uniform-random steps, auto-reset, injected scenes")


if __name__ == "__main__":
    main()
