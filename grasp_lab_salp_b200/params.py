"""SalpParams: every literal the reference hard-codes on the SalpRobotEnv hot path.

ctypes mirror of `struct SalpParams` in include/salp_b200.h.  Defaults follow
make_env() (reference src/train_robot.py:11-21), Robot.__init__ (src/robot.py:261-308),
Nozzle.__init__ (src/robot.py:20-47) and SalpRobotEnv.__init__ (src/salp_robot_env.py:35-47).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

MAX_OBSTACLES = 8
OBS_BASE = 6
NUM_REWARD_TERMS = 8
NUM_EPISODE_METRICS = 20

PRECISION_F64 = 0
PRECISION_MIXED = 1

STEP_AUTORESET = 1
STEP_SORT_BY_K = 2
STEP_PIPELINE = 4
STEP_FUSED = 8
STEP_GENERIC = 16
STEP_CHECK_HANDOFF = 32

RAND_ACTION, RAND_OBSERVATION, RAND_DYNAMICS, RAND_DISTURBANCE, RAND_LATENCY = 1, 2, 4, 8, 16

REWARD_TERM_NAMES = ("rewards/track", "rewards/heading", "rewards/smooth", "rewards/yaw",
                     "rewards/time", "rewards/sideslip", "rewards/obstacle")

# row layout of SalpStepIO.episode_metrics (enum SalpEpisodeMetric)
EPISODE_METRIC_NAMES = (
    "r", "l", "path_length", "direct_distance", "path_efficiency", "final_distance",
    "initial_distance", "avg_compression", "avg_coast_time", "avg_nozzle_angle", "avg_velocity",
    "avg_rewards_track", "avg_rewards_heading", "avg_rewards_smooth", "avg_rewards_yaw",
    "avg_rewards_time", "avg_rewards_sideslip", "avg_rewards_obstacle", "total_substeps",
    "nonfinite")


class SalpParams(C.Structure):
    _fields_ = [
        ("nozzle_length1", C.c_double), ("nozzle_length2", C.c_double), ("nozzle_length3", C.c_double),
        ("nozzle_area", C.c_double), ("nozzle_mass", C.c_double),
        ("nozzle_gamma", C.c_double), ("nozzle_angle_speed", C.c_double),
        ("dry_mass", C.c_double), ("init_length", C.c_double), ("init_width", C.c_double),
        ("max_contraction", C.c_double), ("density", C.c_double), ("dt", C.c_double),
        ("buoy_mass", C.c_double), ("skin_mass", C.c_double), ("tube_mass", C.c_double),
        ("tube_volume", C.c_double),
        ("discharge_coefficient", C.c_double), ("drag_force_ratio", C.c_double),
        ("drag_torque_ratio", C.c_double),
        ("added_mass_force", C.c_double * 3), ("added_mass_rate_force", C.c_double * 3),
        ("added_mass_torque", C.c_double * 3), ("added_mass_rate_torque", C.c_double * 3),
        ("trans_drag_range", C.c_double * 6), ("rot_drag_range", C.c_double * 6),
        ("refill_poly", C.c_double * 3), ("jet_poly", C.c_double * 3),
        ("tank_x_min", C.c_double), ("tank_x_max", C.c_double),
        ("tank_y_min", C.c_double), ("tank_y_max", C.c_double),
        ("target_radius", C.c_double), ("obstacle_radius", C.c_double),
        ("out_of_bounds_distance", C.c_double),
        ("success_bonus", C.c_double), ("out_of_bounds_penalty", C.c_double),
        ("collision_penalty", C.c_double), ("timeout_penalty", C.c_double),
        ("max_cycles", C.c_int32), ("num_obstacles", C.c_int32),
        ("precision", C.c_int32), ("randomization", C.c_int32),
    ]

    def copy(self) -> "SalpParams":
        out = SalpParams()
        C.memmove(C.byref(out), C.byref(self), C.sizeof(SalpParams))
        return out

    @property
    def obs_dim(self) -> int:
        return OBS_BASE + 2 * int(self.num_obstacles)


def fit_timing_polynomials():
    """np.polyfit(deg 2) on the reference's data points -- the same call as
    geometry.py:6-10 (refill) and geometry.py:17-21 (propulsion), so the coefficients carry
    the same last-bit noise as the reference's on this machine."""
    compression = np.array([0.01, 0.02, 0.03, 0.04])
    refill = np.polyfit(compression, np.array([0.4, 1.0, 1.8, 2.2]), 2)
    jet = np.polyfit(compression, np.array([0.1, 0.3, 0.4, 0.5]), 2)
    return refill, jet


def default_params(*, precision: int = PRECISION_MIXED, num_obstacles: int = 2,
                   obstacle_radius: float = 0.2, width: int = 900, height: int = 700,
                   refill_poly=None, jet_poly=None, randomization: int = 0) -> SalpParams:
    if not 0 <= num_obstacles <= MAX_OBSTACLES:
        raise ValueError(f"num_obstacles must be in [0, {MAX_OBSTACLES}]")
    p = SalpParams()
    # Nozzle(length1=0.05, length2=0.05, length3=0.05, area=0.00016, mass=1.0)
    p.nozzle_length1 = p.nozzle_length2 = p.nozzle_length3 = 0.05
    p.nozzle_area = 0.00016
    p.nozzle_mass = 1.0
    p.nozzle_gamma = np.pi / 4
    p.nozzle_angle_speed = 31 * np.pi / 30
    # Robot(dry_mass=1.0, init_length=0.3, init_width=0.15, max_contraction=0.06)
    p.dry_mass = 1.0
    p.init_length = 0.3
    p.init_width = 0.15
    p.max_contraction = 0.06
    p.density = 1000.0
    p.dt = 0.01
    p.buoy_mass = 0.195
    p.skin_mass = 0.145
    p.tube_mass = 0.414
    p.tube_volume = np.pi * (0.058 / 2) ** 2 * 0.15
    p.discharge_coefficient = 0.3
    p.drag_force_ratio = 0.25
    p.drag_torque_ratio = 0.1
    p.added_mass_force[:] = [0.5, 0.6, 0.6]
    p.added_mass_rate_force[:] = [0.2, 0.2, 0.2]
    p.added_mass_torque[:] = [0.3, 0.6, 0.6]
    p.added_mass_rate_torque[:] = [0.2, 0.2, 0.2]
    p.trans_drag_range[:] = [1.5, 2.5, 2.5, 1.5, 2.5, 1.5]
    p.rot_drag_range[:] = [0.1, 0.3, 0.5, 0.2, 0.5, 0.2]
    if refill_poly is None or jet_poly is None:
        r, j = fit_timing_polynomials()
        refill_poly = r if refill_poly is None else refill_poly
        jet_poly = j if jet_poly is None else jet_poly
    p.refill_poly[:] = [float(x) for x in refill_poly]
    p.jet_poly[:] = [float(x) for x in jet_poly]
    # SalpRobotEnv: tank bounds in metres (salp_robot_env.py:471-482), scale = 200 px/m, margin 50 px
    scale, margin = 200.0, 50
    p.tank_x_min = (-width / 2 + margin) / scale
    p.tank_x_max = (width / 2 - margin) / scale
    p.tank_y_min = (-height / 2 + margin) / scale
    p.tank_y_max = (height / 2 - margin) / scale
    p.target_radius = 0.2
    p.obstacle_radius = obstacle_radius
    p.out_of_bounds_distance = 5.0
    p.success_bonus = 500.0
    p.out_of_bounds_penalty = 200.0
    p.collision_penalty = 200.0
    p.timeout_penalty = 50.0
    p.max_cycles = 500
    p.num_obstacles = num_obstacles
    p.precision = precision
    p.randomization = int(randomization)
    return p


# ---- SalpField ids (enum SalpField in include/salp_b200.h) --------------------------------
_F64_NAMES = (
    "vel_x vel_y vel_z angvel_x angvel_y angvel_z euler_x euler_y euler_z posw_x posw_y posw_z "
    "acc_x acc_y acc_z angacc_x angacc_y angacc_z pos_x pos_y pos_z angle_x angle_y angle_z "
    "prevpos_x prevpos_y prevpos_z prevangle_x prevangle_y prevangle_z "
    "length width prev_volume prev_i_x prev_i_y prev_i_z "
    "com_x prev_com_x com_rate_x prev_com_rate_x com_acc_x "
    "nozzle_angle1 nozzle_angle2 prev_dist speed_world "
    "ep_return ep_path_length ep_initial_distance ep_sum_a0 ep_sum_a1 ep_sum_abs_a2 ep_sum_speed "
    "ep_sum_term0 ep_sum_term1 ep_sum_term2 ep_sum_term3 ep_sum_term4 ep_sum_term5 ep_sum_term6 "
    "ep_substeps ou_force_x ou_force_y ou_torque_z").split()
F32_BASE = 1000
I32_BASE = 2000
_F32_NAMES = ["nozzle_yaw", "prev_action0", "prev_action1", "prev_action2", "target_x", "target_y"]
for _i in range(MAX_OBSTACLES):
    _F32_NAMES += [f"obstacle{_i}_x", f"obstacle{_i}_y"]
_I32_NAMES = ["phase", "cycle", "ep_length", "episode_index"]

FIELDS = {}
for _i, _n in enumerate(_F64_NAMES):
    FIELDS[_n] = (_i, np.float64)
for _i, _n in enumerate(_F32_NAMES):
    FIELDS[_n] = (F32_BASE + _i, np.float32)
for _i, _n in enumerate(_I32_NAMES):
    FIELDS[_n] = (I32_BASE + _i, np.int32)
NUM_F64_FIELDS = len(_F64_NAMES)


def field_id(name: str) -> int:
    return FIELDS[name][0]


def field_dtype(name: str):
    return FIELDS[name][1]


# Batches larger than one warp per SM sub-partition (148 SMs x 4 x 32 lanes) step faster in K-sorted
# order (measured crossover, tools/diag_crossover.py: 24576 envs 0.33 ms sorted vs 0.38 ms unsorted;
# at or below 18944 envs every warp has its own scheduler and the 3 sort launches only cost time).
SORT_AUTO_MIN_ENVS = 148 * 4 * 32 + 1


def sort_by_k_auto(num_envs):
    return num_envs >= SORT_AUTO_MIN_ENVS
