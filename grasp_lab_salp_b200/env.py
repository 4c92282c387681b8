"""Gymnasium-style single-env view and constructor-compatible stand-ins for the reference classes.

``make_env()`` of the reference trainers (src/train_robot.py:11-21) reads

    nozzle = Nozzle(length1=0.05, length2=0.05, length3=0.05, area=0.00016, mass=1.0)
    robot = Robot(dry_mass=1.0, init_length=0.3, init_width=0.15, max_contraction=0.06, nozzle=nozzle)
    robot.nozzle.set_angles(angle1=0.0, angle2=0.0)
    robot.set_environment(density=1000)
    env = SalpRobotEnv(render_mode=None, robot=robot)

With ``from grasp_lab_salp_b200.env import Nozzle, Robot, SalpRobotEnv`` those lines run
unchanged: ``Nozzle`` / ``Robot`` only record their constructor arguments, ``SalpRobotEnv`` turns
them into a ``SalpParams`` and drives a 1-env ``SalpBatch`` on the GPU.  reset()/step() have the
reference's signatures and return types (src/salp_robot_env.py:114-155, 196-299).
"""
from __future__ import annotations

import numpy as np

from . import spaces
from .batch import SalpBatch
from .params import EPISODE_METRIC_NAMES, REWARD_TERM_NAMES, SalpParams, default_params

try:  # pragma: no cover - gymnasium is not in the build image
    import gymnasium as _gym
    _EnvBase = _gym.Env
except Exception:
    _EnvBase = object


class Nozzle:
    """Constructor-compatible with reference Nozzle (src/robot.py:20-47)."""

    def __init__(self, length1: float = 0.0, length2: float = 0.0, length3: float = 0.0, area: float = 0.0,
                 mass: float = 0.0):
        self.length1, self.length2, self.length3, self.area, self.mass = length1, length2, length3, area, mass
        self.angle1 = 0.0
        self.angle2 = 0.0
        self.yaw = 0.0

    def set_angles(self, angle1: float, angle2: float):
        self.angle1, self.angle2 = float(angle1), float(angle2)


class Robot:
    """Constructor-compatible with reference Robot (src/robot.py:261-308).  After it is handed to
    SalpRobotEnv its state attributes (position_world, position, velocity, euler_angle,
    angular_velocity, cycle, length, width, state) read through to the GPU columns."""

    def __init__(self, dry_mass: float, init_length: float, init_width: float, max_contraction: float,
                 nozzle: Nozzle):
        self.dry_mass, self.init_length, self.init_width = dry_mass, init_length, init_width
        self.max_contraction = max_contraction
        self.nozzle = nozzle
        self.density = 0.0
        self._batch: SalpBatch | None = None

    def set_environment(self, density: float):
        self.density = density

    def _vec(self, prefix):
        b = self._batch
        return np.array([b.get_state(f"{prefix}_{a}")[0] for a in "xyz"])

    @property
    def position_world(self):
        return self._vec("posw")

    @property
    def position(self):
        return self._vec("pos")

    @property
    def velocity(self):
        return self._vec("vel")

    @property
    def euler_angle(self):
        return self._vec("euler")

    @property
    def angular_velocity(self):
        return self._vec("angvel")

    @property
    def cycle(self):
        return int(self._batch.get_state("cycle")[0])

    @property
    def length(self):
        return float(self._batch.get_state("length")[0])

    @property
    def width(self):
        return float(self._batch.get_state("width")[0])


def params_from_robot(robot: Robot | None, *, width=900, height=700, num_obstacles=2, obstacle_radius=0.2,
                      precision=None) -> SalpParams:
    kw = dict(num_obstacles=num_obstacles, obstacle_radius=obstacle_radius, width=width, height=height)
    if precision is not None:
        kw["precision"] = precision
    p = default_params(**kw)
    if robot is not None:
        p.dry_mass = robot.dry_mass
        p.init_length = robot.init_length
        p.init_width = robot.init_width
        p.max_contraction = robot.max_contraction
        p.density = robot.density
        nz = robot.nozzle
        p.nozzle_length1, p.nozzle_length2, p.nozzle_length3 = nz.length1, nz.length2, nz.length3
        p.nozzle_area, p.nozzle_mass = nz.area, nz.mass
    return p


class SalpRobotEnv(_EnvBase):
    """Single-env gymnasium surface (reference src/salp_robot_env.py:22-299), one GPU env.
    A `gymnasium.Env` subclass wherever gymnasium is importable (so `check_env`, `Monitor` and
    `make_vec_env` accept it); a plain class with the same methods otherwise."""

    metadata = {"render_modes": ["human", "rgb_array"], "render_fps": 60}

    def __init__(self, render_mode=None, width: int = 900, height: int = 700, robot: Robot | None = None,
                 num_obstacles: int = 2, obstacle_radius: float = 0.2, *, seed: int = 0, device: int = 0,
                 precision=None, record: bool = False, _cdll=None):
        if render_mode is not None:
            raise NotImplementedError("rendering is out of scope of the GPU simulator (render_mode must be None)")
        self.render_mode = None
        self.width, self.height = width, height
        self.num_obstacles, self.obstacle_radius = num_obstacles, obstacle_radius
        self.target_radius = 0.2
        self.robot = robot if robot is not None else Robot(1.0, 0.3, 0.15, 0.06, Nozzle(0.05, 0.05, 0.05, 0.00016, 1.0))
        if robot is None:
            self.robot.set_environment(1000)
        self.params = params_from_robot(self.robot, width=width, height=height, num_obstacles=num_obstacles,
                                        obstacle_radius=obstacle_radius, precision=precision)
        self._batch = SalpBatch(1, self.params, seed=seed, device=device, _cdll=_cdll)
        self.robot._batch = self._batch
        if self.robot.nozzle.angle1 or self.robot.nozzle.angle2:
            self._batch.set_state("nozzle_angle1", self.robot.nozzle.angle1)
            self._batch.set_state("nozzle_angle2", self.robot.nozzle.angle2)
        self.action_space = spaces.action_space()
        self.observation_space = spaces.observation_space(num_obstacles)
        self.action = np.zeros(3)
        self.record = bool(record)         # Robot.enable_history_recording(): per-substep histories in `info`
        self.last_history = None
        self.reset()                       # the reference resets in its constructor (salp_robot_env.py:112)

    def enable_history_recording(self):
        self.record = True

    def disable_history_recording(self):
        self.record = False

    # ---- gymnasium API ----
    def reset(self, seed=None, options=None):
        if _EnvBase is not object:
            super().reset(seed=seed)       # gymnasium bookkeeping (np_random); the reference never uses it either
        if seed is not None:
            self.action_space.seed(seed)
        obs = self._batch.reset()
        return obs[0].copy(), {}

    def step(self, action):
        a = np.asarray(action, np.float32).reshape(1, 3)
        self.action = a[0]
        b = self._batch
        hist = b.trace_cycle(0, a[0]) if self.record else None     # (does not advance the env)
        self.last_history = hist
        obs, rew, term, trunc = b.step(a, auto_reset=False, extras=True)
        if hist is None:
            info = {"position_history": [], "length_history": [], "width_history": []}
        else:
            info = {"position_history": hist["position_world"], "length_history": hist["length"],
                    "width_history": hist["width"]}
        for j, k in enumerate(REWARD_TERM_NAMES):
            info[k] = float(b.terms[0, j])
        done, truncated = bool(term[0]), bool(trunc[0])
        if done or truncated:
            for j, k in enumerate(EPISODE_METRIC_NAMES[2:18]):
                info[k] = float(b.metrics[0, 2 + j])
            if b.metrics[0, 19] == 1.0:
                info["numerical_blowup"] = True
        return obs[0].copy(), float(b.terms[0, 7]), done, truncated, info

    def render(self):
        return None

    def close(self):
        self._batch.close()

    # ---- attributes the reference's tools read (watch_model.py:72, salp_robot_env.py:1502) ----
    @property
    def target_point(self):
        return np.array([self._batch.get_state("target_x")[0], self._batch.get_state("target_y")[0]], np.float32)

    @property
    def obstacles(self):
        return [np.array([self._batch.get_state(f"obstacle{k}_x")[0], self._batch.get_state(f"obstacle{k}_y")[0]],
                         np.float32) for k in range(self.num_obstacles)]

    def set_scene(self, target, obstacles):
        """Test hook: the next reset() uses this target/obstacle set instead of a sampled one."""
        t = np.asarray(target, np.float32).reshape(1, 1, 2)
        o = np.asarray(obstacles, np.float32).reshape(1, 1, self.num_obstacles, 2)
        self._batch.set_scene_pool(t, o)


SalpCudaEnv = SalpRobotEnv
