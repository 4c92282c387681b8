"""grasp_lab_salp_b200 -- B200-native batched SALP robot simulator."""
from .params import (SalpParams, default_params, PRECISION_F64, PRECISION_MIXED,  # noqa: F401
                     STEP_AUTORESET, STEP_SORT_BY_K)
from .batch import SalpBatch  # noqa: F401
from ._lib import SalpError  # noqa: F401
from .vec_env import SalpCudaVecEnv  # noqa: F401
from .env import Nozzle, Robot, SalpCudaEnv, SalpRobotEnv  # noqa: F401

__all__ = ["SalpParams", "default_params", "PRECISION_F64", "PRECISION_MIXED",
           "STEP_AUTORESET", "STEP_SORT_BY_K", "SalpBatch", "SalpError", "SalpCudaVecEnv", "SalpCudaEnv",
           "SalpRobotEnv", "Robot", "Nozzle"]
