"""Action / observation spaces of SalpRobotEnv (reference src/salp_robot_env.py:63-75).

gymnasium's Box is used when gymnasium is importable (so SB3 type checks pass); otherwise a
minimal stand-in with the same attributes (low, high, shape, dtype, sample, contains, seed).
"""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - depends on the environment
    from gymnasium.spaces import Box as _GymBox
except Exception:  # gymnasium is not installed in the build image
    _GymBox = None


class _Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        if shape is None:
            shape = np.shape(low)
        self.shape = tuple(shape)
        self.low = np.broadcast_to(np.asarray(low, self.dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, self.dtype), self.shape).copy()
        self._rng = np.random.default_rng()

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)
        return [seed]

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return self._rng.uniform(lo, hi).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"


Box = _GymBox or _Box


def action_space():
    """[inhale_control 0..1, coast_time 0..1, nozzle_direction -1..1]  (salp_robot_env.py:63-67)"""
    return Box(low=np.array([0.0, 0.0, -1.0], np.float32), high=np.array([1.0, 1.0, 1.0], np.float32),
               dtype=np.float32)


def observation_space(num_obstacles: int):
    """6 + 2 * num_obstacles unbounded float32 (salp_robot_env.py:70-75)"""
    return Box(low=-np.inf, high=np.inf, shape=(6 + 2 * num_obstacles,), dtype=np.float32)
