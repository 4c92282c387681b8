#!/usr/bin/env python
"""Generate tests/golden/ref_*.npz from the LIVE, UNMODIFIED Python reference.

Runs only in the build container (needs /root/reference or $SALP_REF_DIR, numpy, numba).
The reference is imported read-only through oracle/ref_harness.py; nothing is copied.
Each trace drives SalpRobotEnv exactly like an SB3 VecEnv worker would (step, and on
done/truncated: reset), with the scene that reset() sampled from the global np.random
replaced by a recorded one (ref_harness.inject_scene) so any other implementation can
be driven with the identical targets/obstacles.

    python tools/gen_golden.py            # writes tests/golden/ref_{fixed10,edge,random,clipped}.npz

Recorded per (env, step): action, obs (of the finished step, i.e. the terminal observation
when the episode ends), post-reset obs, reward, 7 reward components, terminated, truncated,
K (substeps run), Robot.cycle, Robot.state, and a state vector (see STATE_NAMES).
"""
from __future__ import annotations

import os
import sys
from multiprocessing import Pool

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402

STATE_NAMES = ["posw_x", "posw_y", "posw_z", "vel_x", "vel_y", "vel_z", "euler_x", "euler_y", "euler_z",
               "angvel_x", "angvel_y", "angvel_z", "length", "width", "volume",
               "nozzle_angle1", "nozzle_angle2", "turn_time", "refill_time", "jet_time", "prev_dist",
               "pos_x", "pos_y", "pos_z", "angle_x", "angle_y", "angle_z",
               "acc_x", "acc_y", "acc_z", "angacc_x", "angacc_y", "angacc_z", "com_x", "total_cycle_time"]
TERM_KEYS = ["rewards/track", "rewards/heading", "rewards/smooth", "rewards/yaw", "rewards/time",
             "rewards/sideslip", "rewards/obstacle"]
METRIC_KEYS = ["path_length", "direct_distance", "path_efficiency", "final_distance", "initial_distance",
               "avg_compression", "avg_coast_time", "avg_nozzle_angle", "avg_velocity",
               "avg_rewards_track", "avg_rewards_heading", "avg_rewards_smooth", "avg_rewards_yaw",
               "avg_rewards_time", "avg_rewards_sideslip", "avg_rewards_obstacle"]

FIXED10 = np.array([  # the reference's own fixed input set, salp_robot_env.py:1568-1579
    [0.695722, 0.01922786, -0.06692487], [0.2808507, 0.8017318, 0.87773895],
    [0.57452214, 0.11145315, -0.82465506], [0.32618135, 0.11088043, 0.88842094],
    [0.17267734, 0.6958977, -0.9337022], [0.49285844, 0.2883283, 0.81122017],
    [0.34796143, 0.35572827, -0.8472595], [0.49369425, 0.27951986, 0.8069289],
    [0.37975544, 0.338947, -0.8655774], [0.4979022, 0.23918751, 0.7962456]], dtype=np.float32)


def sample_scenes(rng, n_envs, P, num_obstacles=2):
    """Host-side scene sampler with the reference's rejection rule (salp_robot_env.py:535-559)."""
    targets = np.zeros((n_envs, P, 2), np.float32)
    obstacles = np.zeros((n_envs, P, num_obstacles, 2), np.float32)
    for i in range(n_envs):
        for s in range(P):
            t = np.array([rng.uniform(-2, 2), rng.uniform(-1.5, 1.5)]).astype(np.float32)
            obs = []
            while len(obs) < num_obstacles:
                pos = np.array([rng.uniform(-2, 2), rng.uniform(-1.5, 1.5)], dtype=np.float32)
                if (np.linalg.norm(pos) > 0.5 and np.linalg.norm(pos - t) > 0.5
                        and not any(np.linalg.norm(pos - o) < 0.5 for o in obs)):
                    obs.append(pos)
            targets[i, s] = t
            obstacles[i, s] = np.array(obs)
    return targets, obstacles


def _state_vector(env, total):
    r = env.robot
    return np.array([*r.position_world, *r.velocity, *r.euler_angle, *r.angular_velocity,
                     r.length, r.width, r.volume, r.nozzle.angle1, r.nozzle.angle2, r.nozzle.turn_time,
                     r.refill_time, r.jet_time, env.prev_dist, *r.position, *r.angle,
                     *r.acceleration, *r.angular_acceleration, r.center_of_mass[0], total], dtype=np.float64)


def run_env(args):
    """One reference env, T steps, SB3-worker auto-reset with injected scenes."""
    actions, targets, obstacles = args
    T = actions.shape[0]
    P = targets.shape[0]
    env = rh.make_env()
    counter = {"k": 0}
    orig_step = env.robot.step

    def counting_step():
        counter["k"] += 1
        orig_step()

    env.robot.step = counting_step     # instance attribute: the `while` loop of robot.py:756-757 calls it
    env.reset()
    episode = 0
    obs0 = rh.inject_scene(env, targets[0], obstacles[0])
    D = obs0.shape[0]
    out = dict(obs=np.zeros((T, D), np.float32), reset_obs=np.zeros((T, D), np.float32),
               reward=np.zeros(T), terms=np.zeros((T, 7)), terminated=np.zeros(T, np.uint8),
               truncated=np.zeros(T, np.uint8), K=np.zeros(T, np.int32), cycle=np.zeros(T, np.int32),
               phase=np.zeros(T, np.int32), state=np.zeros((T, len(STATE_NAMES))),
               metrics=np.full((T, len(METRIC_KEYS)), np.nan), first_obs=obs0)
    for t in range(T):
        counter["k"] = 0
        obs, rew, done, trunc, info = env.step(actions[t].copy())
        r = env.robot
        total = max(r.refill_time, r.nozzle.turn_time) + r.jet_time + r.coast_time
        out["obs"][t] = obs
        out["reward"][t] = rew
        out["terms"][t] = [info[k] for k in TERM_KEYS]
        out["terminated"][t] = done
        out["truncated"][t] = trunc
        out["K"][t] = counter["k"]
        out["cycle"][t] = r.cycle
        out["phase"][t] = r.state.value
        out["state"][t] = _state_vector(env, float(total))
        if done or trunc:
            out["metrics"][t] = [info.get(k, np.nan) for k in METRIC_KEYS]
            episode += 1
            env.reset()
            out["reset_obs"][t] = rh.inject_scene(env, targets[episode % P], obstacles[episode % P])
        else:
            out["reset_obs"][t] = obs
    return out


def run_blowup_case(args):
    """One cycle from rest with given carried nozzle angles; does the reference raise?"""
    action, angle1, angle2 = args
    env = rh.make_env()
    env.reset()
    rh.inject_scene(env, [1.8, 1.2], [[-1.5, -1.0], [1.5, -1.0]])
    env.robot.nozzle.angle1 = float(angle1)
    env.robot.nozzle.angle2 = float(angle2)
    counter = {"k": 0}
    orig_step = env.robot.step

    def counting_step():
        counter["k"] += 1
        orig_step()

    env.robot.step = counting_step
    raised = 0
    try:
        env.step(np.asarray(action, np.float32).copy())
    except np.linalg.LinAlgError:
        raised = 1
    r = env.robot
    return raised, counter["k"], np.array([*r.position_world[:2], r.euler_angle[2]], np.float64)


RAND_ACTIONS = np.array([[0.6, 0.2, 0.3], [0.5, 0.3, -0.4], [0.7, 0.1, 0.2]], np.float32)


def run_rand_case(args):
    """3 fixed cycles from rest with ONE of the reference's robustness switches on; global
    np.random seeded per sample.  Returns final (x, y, yaw, vx, vy, wz), last obs[:6], yaw command."""
    mode, seed = args
    np.random.seed(seed)
    env = rh.make_env()
    if mode == "dynamics":
        env.robot.enable_dynamic_randomization()
    elif mode == "disturbance":
        env.robot.enable_disturbances()
    elif mode == "action":
        env.enable_action_randomization()
    elif mode == "observation":
        env.enable_observation_randomization()
    env.reset()
    rh.inject_scene(env, [1.8, 1.2], [[-1.5, -1.0], [1.5, -1.0]])
    obs = None
    for a in RAND_ACTIONS:
        obs, rew, done, trunc, info = env.step(a.copy())
    r = env.robot
    return np.array([r.position_world[0], r.position_world[1], r.euler_angle[2], r.velocity[0], r.velocity[1],
                     r.angular_velocity[2], *obs[:6], float(r.nozzle.yaw)], np.float64)


def gather(results):
    keys = results[0].keys()
    return {k: np.stack([r[k] for r in results]) for k in keys}


def write(name, actions, targets, obstacles, note):
    """actions [N,T,3] f32, targets [N,P,2], obstacles [N,P,2,2]."""
    n = actions.shape[0]
    with Pool(min(8, n)) as pool:
        res = pool.map(run_env, [(actions[i], targets[i], obstacles[i]) for i in range(n)])
    g = gather(res)
    m = rh.load()
    path = os.path.join(ROOT, "tests", "golden", name)
    np.savez_compressed(
        path, actions=actions, targets=targets, obstacles=obstacles,
        refill_poly=m.geometry.fit_compression_refill_time_relation_jit(),
        jet_poly=m.geometry.fit_compression_propulsion_time_relation_jit(),
        state_names=np.array(STATE_NAMES), metric_keys=np.array(METRIC_KEYS), note=np.array(note), **g)
    print(f"{name}: {n} envs x {actions.shape[1]} steps, mean K {g['K'].mean():.1f}, "
          f"episodes ended {int((g['terminated'] | g['truncated']).sum())}")


def main():
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    which = sys.argv[1:] or ["fixed10", "edge", "random", "clipped", "blowup", "randstats"]
    rng = np.random.default_rng(20261018)

    if "fixed10" in which:
        # SURVEY.md 8c seed KAT scene + the reference's own 10 fixed actions, twice over (20 steps)
        t = np.array([[[1.5, -0.75]]], np.float32)
        o = np.array([[[[0.8, 0.6], [-1.0, -1.0]]]], np.float32)
        write("ref_fixed10.npz", np.concatenate([FIXED10, FIXED10])[None], t, o,
              "salp_robot_env.py:1568-1579 actions x2, scene of SURVEY 8c")

    if "edge" in which:
        # edge actions, each sequence from rest (SURVEY.md 8c "Edge KATs") + continuation steps
        seqs = [
            [(0, 0, 0), (0, 0, 0), (1, 1, 1), (0, 0, 0)],
            [(1, 1, 1), (1, 0, -1), (0, 1, 0), (1, 1, -1)],
            [(1, 0, -1), (1, 0, 1), (1, 0, -1), (1, 0, 1)],
            [(0.088, 0, 0.5), (0.089, 0, -0.5), (0.0885, 0.001, 0.0), (0.5, 0, 0.0)],
            [(0.5, 0.5, 0.0), (0.5, 0.5, 1e-4), (0.5, 0.5, -1e-4), (0.5, 0.5, 2e-4)],
            [(1, 0, 0.01), (1, 0, -0.01), (0.3, 0, 0.02), (0.05, 0.02, 1.0)],
            [(0.09, 0, 0), (0.2, 0, 0), (0.0, 0.3, 1.0), (0.0, 0.0, -1.0)],
            [(1, 0, 0), (1, 0, 0), (1, 0, 0), (1, 0, 0)],
        ]
        a = np.array(seqs, np.float32)
        t, o = sample_scenes(rng, len(seqs), 4)
        t[:, 0] = [1.8, 1.2]         # far target first so edge sequences are not cut short
        write("ref_edge.npz", a, t, o, "edge actions: zeros, ones, K=0, negative refill/jet, yaw~0, coast=0")

    if "random" in which:
        n, T = 24, 30
        a = np.stack([rng.uniform([0, 0, -1], [1, 1, 1], size=(T, 3)) for _ in range(n)]).astype(np.float32)
        t, o = sample_scenes(rng, n, 8)
        write("ref_random.npz", a, t, o, "uniform Box actions (SURVEY 8d input A)")

    if "clipped" in which:
        # random-init-policy-like actions: N(0,1) clipped to the Box (SURVEY 8d input B)
        n, T = 8, 30
        a = np.clip(rng.normal(size=(n, T, 3)), [0, 0, -1], [1, 1, 1]).astype(np.float32)
        t, o = sample_scenes(rng, n, 8)
        write("ref_clipped.npz", a, t, o, "N(0,1) actions clipped to the Box (SURVEY 8d input B)")


    if "blowup" in which:
        # contractions whose jet_time is a fraction of one substep: the reference's integrator can
        # diverge and np.linalg.solve then raises LinAlgError (dynamics.py:6-10)
        n = 512
        a = rng.uniform([0.0885, 0.0, -1.0], [0.0945, 1.0, 1.0], size=(n, 3)).astype(np.float32)
        ang1 = rng.uniform(-np.pi, np.pi, n)
        ang2 = rng.uniform(0.0, np.pi, n)
        with Pool(8) as pool:
            res = pool.map(run_blowup_case, [(a[i], ang1[i], ang2[i]) for i in range(n)])
        m = rh.load()
        np.savez_compressed(
            os.path.join(ROOT, "tests", "golden", "ref_blowup.npz"), actions=a, angle1=ang1, angle2=ang2,
            raised=np.array([r[0] for r in res], np.uint8), substeps_run=np.array([r[1] for r in res], np.int32),
            final=np.stack([r[2] for r in res]),
            refill_poly=m.geometry.fit_compression_refill_time_relation_jit(),
            jet_poly=m.geometry.fit_compression_propulsion_time_relation_jit(),
            note=np.array("one cycle from rest, carried nozzle angles injected; raised = reference threw LinAlgError"))
        print(f"ref_blowup.npz: {n} cases, reference raised in {sum(r[0] for r in res)}")


    if "randstats" in which:
        # distributions of the outcome of 3 fixed cycles under each default-off robustness switch
        n = 320
        out = {}
        with Pool(8) as pool:
            for mode in ("none", "dynamics", "disturbance", "action", "observation"):
                res = pool.map(run_rand_case, [(mode, 1000 + i) for i in range(n if mode != "none" else 1)])
                out[mode] = np.stack(res)
                print(mode, "mean", out[mode].mean(0)[:6], "std", out[mode].std(0)[:6])
        m = rh.load()
        np.savez_compressed(
            os.path.join(ROOT, "tests", "golden", "ref_randstats.npz"), actions=RAND_ACTIONS,
            columns=np.array(["x", "y", "yaw", "vx", "vy", "wz", "obs0", "obs1", "obs2", "obs3", "obs4", "obs5", "nozzle_yaw"]),
            refill_poly=m.geometry.fit_compression_refill_time_relation_jit(),
            jet_poly=m.geometry.fit_compression_propulsion_time_relation_jit(),
            note=np.array("3 fixed cycles from rest, scene target (1.8,1.2); one robustness switch per set; 320 samples each"),
            **{f"samples_{k}": v for k, v in out.items()})


if __name__ == "__main__":
    main()
