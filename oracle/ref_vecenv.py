"""TEST / BENCH INFRASTRUCTURE ONLY -- the reference's CPU vectorised path, unmodified.

`make_vec_env(make_env, n_envs, vec_env_cls=SubprocVecEnv)` of the reference's trainer
(src/train_robot.py:25-26) with stable-baselines3 absent from this image: one OS process per env,
pipe IPC, lock-step `step_async` / `step_wait`, and the worker-side auto-reset of SB3's
`_worker` (reset on done, `terminal_observation` kept).  Every worker runs the reference's own
`SalpRobotEnv` (unmodified source, loaded by oracle/ref_harness.py from $SALP_REF_DIR,
baseline/_ref/src or /root/reference/src).  Used by bench.py's CPU legs only.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np

from . import ref_harness as rh


def _worker(conn, seed):
    np.random.seed(seed)                      # reset() draws targets / obstacles from the GLOBAL np.random
    env = rh.make_env()
    substeps = [0]
    orig = env.robot.step

    def counting_step():                      # instance attribute: robot.py:756-757 calls self.step()
        substeps[0] += 1
        orig()

    env.robot.step = counting_step
    obs, _ = env.reset()
    try:
        while True:
            cmd, data = conn.recv()
            if cmd == "step":
                substeps[0] = 0
                try:
                    obs, rew, done, trunc, info = env.step(data)
                except np.linalg.LinAlgError:         # the reference's integrator blew up: episode over
                    obs, rew, done, trunc, info = obs, -200.0, False, True, {}
                    env.robot.reset()
                if done or trunc:
                    info = dict(terminal_observation=obs)
                    obs, _ = env.reset()
                else:
                    info = {}
                conn.send((obs, rew, done or trunc, info, substeps[0]))
            elif cmd == "reset":
                obs, _ = env.reset()
                conn.send(obs)
            else:
                break
    finally:
        conn.close()


class RefSubprocVecEnv:
    """n_envs workers, one reference SalpRobotEnv each (SubprocVecEnv semantics)."""

    def __init__(self, n_envs: int, seed: int = 0):
        if not rh.available():
            raise RuntimeError("reference sources (or numba) not available")
        # JIT-compile the reference's numba kernels once in the parent: forked workers inherit them
        warm = rh.make_env()
        warm.reset()
        warm.step(np.array([0.5, 0.1, 0.3], np.float32))
        ctx = mp.get_context("fork")
        self.n = n_envs
        self.conns, self.procs = [], []
        for i in range(n_envs):
            a, b = ctx.Pipe()
            p = ctx.Process(target=_worker, args=(b, seed + 1000 * i + 1), daemon=True)
            p.start()
            b.close()
            self.conns.append(a)
            self.procs.append(p)

    def step(self, actions: np.ndarray):
        for c, a in zip(self.conns, actions):          # step_async
            c.send(("step", np.asarray(a, np.float32)))
        res = [c.recv() for c in self.conns]           # step_wait
        obs = np.stack([r[0] for r in res])
        rew = np.array([r[1] for r in res], np.float32)
        done = np.array([r[2] for r in res])
        return obs, rew, done, [r[3] for r in res], int(sum(r[4] for r in res))

    def close(self):
        for c in self.conns:
            try:
                c.send(("close", None))
            except Exception:
                pass
        for p in self.procs:
            p.join(timeout=2)
            if p.is_alive():
                p.terminate()


def time_reference(n_envs: int, vec_steps: int, warmup: int, seed: int = 0):
    """(env_steps, substeps, seconds) of `vec_steps` lock-step vector steps under uniform-random actions."""
    venv = RefSubprocVecEnv(n_envs, seed)
    rng = np.random.default_rng(seed)
    try:
        for _ in range(warmup):
            venv.step(rng.uniform([0, 0, -1], [1, 1, 1], size=(n_envs, 3)).astype(np.float32))
        sub = 0
        t0 = time.perf_counter()
        for _ in range(vec_steps):
            sub += venv.step(rng.uniform([0, 0, -1], [1, 1, 1], size=(n_envs, 3)).astype(np.float32))[4]
        dt = time.perf_counter() - t0
    finally:
        venv.close()
    return n_envs * vec_steps, sub, dt


def manifest():
    """sha256 of the reference files in use (bench.py prints it: evidence that they are unmodified)."""
    import hashlib
    d = rh.reference_dir()
    out = {}
    for f in ("dynamics.py", "geometry.py", "robot.py", "salp_robot_env.py"):
        with open(os.path.join(d, f), "rb") as fh:
            out[f] = hashlib.sha256(fh.read()).hexdigest()[:16]
    return {"dir": d, "sha256_16": out}
