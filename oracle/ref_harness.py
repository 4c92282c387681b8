"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* Python reference.

Imports /root/reference/src/{dynamics,geometry,robot,salp_robot_env}.py read-only
 with throw-away stub modules for the packages that are not
installed in this image (gymnasium, pygame, matplotlib, PIL).  It exists so that

  * tests/golden/*.npz can be (re)generated from the live reference
    (tools/gen_golden.py), and
  * oracle/salp_oracle.c (the C restatement) can be pinned against the reference.

The reference is looked for in $SALP_REF_DIR, then baseline/_ref/src (staged by
tools/stage_reference.py: git-ignored, but it travels to the GPU box with the gpurun snapshot, so
bench.py can time the real reference there), then /root/reference/src (build container only).
Nothing under grasp_lab_salp_b200/ imports this file; of what runs on the GPU box only bench.py's
CPU legs use it (pytest -m gpu and smoke() do not).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

REF_ENV_VAR = "SALP_REF_DIR"
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEARCH = (os.path.join(_REPO, "baseline", "_ref", "src"), "/root/reference/src")


def reference_dir() -> str | None:
    cands = ([os.environ[REF_ENV_VAR]] if os.environ.get(REF_ENV_VAR) else []) + list(SEARCH)
    for d in cands:
        if os.path.isfile(os.path.join(d, "salp_robot_env.py")):
            return d
    return None


def available() -> bool:
    if reference_dir() is None:
        return False
    try:
        import numba  # noqa: F401
    except Exception:
        return False
    return True


def _install_stubs() -> None:
    """Dummy gymnasium / pygame / matplotlib / PIL (none is used on the hot path)."""
    if "gymnasium" not in sys.modules:
        gym = types.ModuleType("gymnasium")

        class Env:  # gymnasium.Env: only reset(seed=) is reached (salp_robot_env.py:115)
            def reset(self, seed=None, options=None):
                self.np_random = np.random.default_rng(seed)

        class Box:  # gymnasium.spaces.Box: constructor + sample only
            def __init__(self, low, high, shape=None, dtype=np.float32):
                self.low = np.asarray(low, dtype=dtype)
                self.high = np.asarray(high, dtype=dtype)
                self.dtype = dtype
                self.shape = self.low.shape
                self._rng = np.random.default_rng()

            def seed(self, s=None):
                self._rng = np.random.default_rng(s)

            def sample(self):
                return self._rng.uniform(self.low, self.high).astype(self.dtype)

        spaces = types.ModuleType("gymnasium.spaces")
        spaces.Box = Box
        gym.Env = Env
        gym.spaces = spaces
        sys.modules["gymnasium"] = gym
        sys.modules["gymnasium.spaces"] = spaces
    for name in ("pygame", "matplotlib", "matplotlib.pyplot", "PIL", "PIL.Image"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if not hasattr(sys.modules["PIL"], "Image"):
        sys.modules["PIL"].Image = sys.modules["PIL.Image"]


_loaded = None


def load():
    """Returns the namespace (robot, salp_robot_env, dynamics, geometry modules)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    d = reference_dir()
    if d is None:
        raise RuntimeError("reference sources not found (set $SALP_REF_DIR)")
    # cache=True wants to write next to the (read-only) source
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/salp_numba_cache")
    _install_stubs()
    if d not in sys.path:
        sys.path.insert(0, d)
    import dynamics
    import geometry
    import robot
    import salp_robot_env

    _loaded = types.SimpleNamespace(dynamics=dynamics, geometry=geometry, robot=robot,
                                    env=salp_robot_env)
    return _loaded


def make_env():
    """Same construction as the reference trainers (train_robot.py:11-21)."""
    m = load()
    nozzle = m.robot.Nozzle(length1=0.05, length2=0.05, length3=0.05, area=0.00016, mass=1.0)
    rob = m.robot.Robot(dry_mass=1.0, init_length=0.3, init_width=0.15, max_contraction=0.06,
                        nozzle=nozzle)
    rob.nozzle.set_angles(angle1=0.0, angle2=0.0)
    rob.set_environment(density=1000)
    return m.env.SalpRobotEnv(render_mode=None, robot=rob)


def inject_scene(env, target, obstacles) -> np.ndarray:
    """Overwrite what reset() sampled from the global np.random (salp_robot_env.py:118-153)
    with a given target/obstacle set, recompute the values derived from them, and
    return the post-reset observation."""
    env.target_point = np.asarray(target, dtype=np.float32)
    env.obstacles = [np.asarray(o, dtype=np.float32) for o in obstacles]
    env.prev_dist = np.linalg.norm(env.robot.position_world[0:-1] - env.target_point)
    env.episode_distances_to_target = [env.prev_dist]
    env.initial_target_distance = env.prev_dist
    return env._get_observation()
