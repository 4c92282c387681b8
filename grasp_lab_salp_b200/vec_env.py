"""SalpCudaVecEnv: the stable-baselines3 VecEnv surface over one SalpBatch (one GPU).

Replaces ``make_vec_env(make_env, n_envs, vec_env_cls=SubprocVecEnv | DummyVecEnv)`` of the
reference trainers (src/train_robot.py:26, src/train_robot_recurrent_ppo.py:63-65) together
with the ``Monitor`` wrapper make_vec_env adds: same ``reset()`` / ``step_async()`` /
``step_wait()`` contract (obs float32 [N,D], rewards float32 [N], dones bool [N], infos list of
dict), same auto-reset semantics (on done: ``info["terminal_observation"]``,
``info["TimeLimit.truncated"]``, returned obs is the post-reset one), same Monitor
``info["episode"] = {"r", "l", "t"}``, and the env's own info keys (7 ``rewards/*`` every step,
the episode metrics of salp_robot_env.py:399-447 on episode end).

If stable_baselines3 is importable the class derives from its ``VecEnv`` ABC, so SB3 algorithms
accept it as is; otherwise (this build image has no SB3) it is duck-typed to the same surface.
"""
from __future__ import annotations

import time
from typing import Any, Sequence

import numpy as np

from . import spaces
from .batch import SalpBatch
from .params import EPISODE_METRIC_NAMES, REWARD_TERM_NAMES, SalpParams, default_params, sort_by_k_auto

try:  # pragma: no cover - depends on the environment
    from stable_baselines3.common.vec_env import VecEnv as _SB3VecEnv
except Exception:
    _SB3VecEnv = None

_EMPTY_HISTORY_KEYS = ("position_history", "length_history", "width_history")   # salp_robot_env.py:279-284
_ENV_METRIC_KEYS = EPISODE_METRIC_NAMES[2:18]     # path_length ... avg_rewards_obstacle


class _VecEnvSurface:
    """Everything SB3's VecEnv ABC requires, written once; mixed into the right base below."""

    metadata = {"render_modes": []}

    def _setup(self, num_envs: int, params: SalpParams | None, seed: int, device: int, env_id_offset: int,
               sort_by_k, info_mode: str, _cdll=None):
        self.params = (params or default_params()).copy()
        self.batch = SalpBatch(num_envs, self.params, seed=seed, env_id_offset=env_id_offset, device=device,
                               _cdll=_cdll)
        self.num_envs = int(num_envs)
        self.observation_space = spaces.observation_space(self.params.num_obstacles)
        self.action_space = spaces.action_space()
        self.render_mode = None
        self.reset_infos = [{} for _ in range(self.num_envs)]
        self._actions = None
        self._seed = seed
        self._t_start = time.time()
        if sort_by_k == "auto":
            sort_by_k = sort_by_k_auto(self.num_envs)
        self._sort = bool(sort_by_k)
        if info_mode not in ("full", "lazy", "auto"):
            raise ValueError("info_mode must be 'full', 'lazy' or 'auto'")
        self._full_infos = info_mode == "full" or (info_mode == "auto" and self.num_envs <= 64)
        self._attrs: dict = {}        # attributes set through set_attr(): name -> {env index: value}

    # ---- VecEnv API ----
    def reset(self):
        self.reset_infos = [{} for _ in range(self.num_envs)]
        return self.batch.reset().copy()

    def step_async(self, actions):
        a = np.asarray(actions, dtype=np.float32)
        if a.shape != (self.num_envs, 3):
            raise ValueError(f"actions must have shape ({self.num_envs}, 3), got {a.shape}")
        self._actions = a

    def step_wait(self):
        if self._actions is None:
            raise RuntimeError("step_wait() called without step_async()")
        b = self.batch
        obs, rew, term, trunc = b.step(self._actions, auto_reset=True, sort_by_k=self._sort, extras=True)
        self._actions = None
        dones = (term | trunc).astype(bool)
        infos = self._make_infos(dones, term, trunc)
        return obs.copy(), rew.copy(), dones, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def _make_infos(self, dones, term, trunc):
        b = self.batch
        n = self.num_envs
        if self._full_infos:
            infos = []
            for i in range(n):
                d = {k: [] for k in _EMPTY_HISTORY_KEYS}
                for j, k in enumerate(REWARD_TERM_NAMES):
                    d[k] = float(b.terms[i, j])
                infos.append(d)
        else:
            shared: dict = {}
            infos = [shared] * n
        for i in np.flatnonzero(dones):
            d = infos[i] if self._full_infos else {REWARD_TERM_NAMES[j]: float(b.terms[i, j]) for j in range(7)}
            m = b.metrics[i]
            for j, k in enumerate(_ENV_METRIC_KEYS):
                d[k] = float(m[2 + j])
            d["episode"] = {"r": float(m[0]), "l": int(m[1]), "t": round(time.time() - self._t_start, 6)}
            d["terminal_observation"] = b.terminal_obs[i].copy()
            d["TimeLimit.truncated"] = bool(trunc[i] and not term[i])
            if m[19] == 1.0:
                d["numerical_blowup"] = True      # the reference would have raised LinAlgError here
            infos[i] = d
        return infos

    def close(self):
        self.batch.close()

    def seed(self, seed=None):
        """SB3 semantics: returns one seed per env.  The scene streams are keyed by (seed, global
        env id, episode); re-seeding takes effect for handles created afterwards."""
        self._seed = seed
        return [None if seed is None else seed + i for i in range(self.num_envs)]

    def _indices(self, indices) -> Sequence[int]:
        if indices is None:
            return range(self.num_envs)
        if isinstance(indices, int):
            return [indices]
        return indices

    _STATE_ATTRS = ("target_point", "obstacles")

    def get_attr(self, attr_name: str, indices=None):
        idx = list(self._indices(indices))
        if attr_name in self._attrs:
            own = self._attrs[attr_name]
            if all(i in own for i in idx):
                return [own[i] for i in idx]
        if attr_name == "render_mode":
            return [None for _ in idx]
        if attr_name == "target_point":
            tx, ty = self.batch.get_state("target_x"), self.batch.get_state("target_y")
            return [np.array([tx[i], ty[i]], np.float32) for i in idx]
        if attr_name == "obstacles":
            cols = [(self.batch.get_state(f"obstacle{k}_x"), self.batch.get_state(f"obstacle{k}_y"))
                    for k in range(self.params.num_obstacles)]
            return [[np.array([cx[i], cy[i]], np.float32) for cx, cy in cols] for i in idx]
        if attr_name in self._attrs:
            own = self._attrs[attr_name]
            return [own.get(i) for i in idx]
        if hasattr(self, attr_name):
            v = getattr(self, attr_name)
            return [v for _ in idx]
        raise AttributeError(attr_name)

    def set_attr(self, attr_name: str, value: Any, indices=None):
        """SB3 semantics: set `attr_name` of the selected envs.  `target_point` / `obstacles` write the
        simulator's scene columns (what the reference's attributes of that name hold); any other name
        is kept per env and handed back by get_attr()."""
        idx = list(self._indices(indices))
        if attr_name == "target_point":
            t = np.asarray(value, np.float32).reshape(2)
            for a, col in zip(t, ("target_x", "target_y")):
                v = self.batch.get_state(col)
                v[idx] = a
                self.batch.set_state(col, v)
            return
        if attr_name == "obstacles":
            o = np.asarray(value, np.float32).reshape(self.params.num_obstacles, 2)
            for k in range(self.params.num_obstacles):
                for a, ax in zip(o[k], "xy"):
                    v = self.batch.get_state(f"obstacle{k}_{ax}")
                    v[idx] = a
                    self.batch.set_state(f"obstacle{k}_{ax}", v)
            return
        own = self._attrs.setdefault(attr_name, {})
        for i in idx:
            own[i] = value

    def env_method(self, method_name: str, *args, indices=None, **kwargs):
        """SB3 semantics: call a method of the selected envs, one result per env.  The envs are rows of
        one batch, so only the reference env's state-free or reset-like methods exist."""
        idx = list(self._indices(indices))
        if method_name == "reset":
            mask = np.zeros(self.num_envs, np.uint8)
            mask[idx] = 1
            obs = self.batch.reset(mask)
            return [(obs[i].copy(), {}) for i in idx]
        if method_name in ("render", "close", "enable_action_randomization", "enable_observation_randomization",
                           "enable_latency"):
            if method_name.startswith("enable_"):
                raise AttributeError(f"{method_name}: robustness switches are construction-time parameters here "
                                     "(SalpParams.randomization), not per-env calls")
            return [None for _ in idx]
        if method_name == "get_wrapper_attr":
            return self.get_attr(args[0], indices)
        raise AttributeError(f"SalpCudaVecEnv has no per-env method {method_name!r} (envs are rows of one batch)")

    def env_is_wrapped(self, wrapper_class, indices=None):
        return [False for _ in self._indices(indices)]

    def get_images(self):
        return [None for _ in range(self.num_envs)]

    def render(self, mode=None):
        return None


if _SB3VecEnv is not None:  # pragma: no cover - SB3 is not in the build image
    class SalpCudaVecEnv(_VecEnvSurface, _SB3VecEnv):
        def __init__(self, num_envs: int, params: SalpParams | None = None, seed: int = 0, device: int = 0,
                     env_id_offset: int = 0, sort_by_k="auto", info_mode: str = "auto", _cdll=None):
            self._setup(num_envs, params, seed, device, env_id_offset, sort_by_k, info_mode, _cdll)
            _SB3VecEnv.__init__(self, self.num_envs, self.observation_space, self.action_space)
else:
    class SalpCudaVecEnv(_VecEnvSurface):
        def __init__(self, num_envs: int, params: SalpParams | None = None, seed: int = 0, device: int = 0,
                     env_id_offset: int = 0, sort_by_k="auto", info_mode: str = "auto", _cdll=None):
            self._setup(num_envs, params, seed, device, env_id_offset, sort_by_k, info_mode, _cdll)
