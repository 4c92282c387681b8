"""TEST INFRASTRUCTURE ONLY: SalpBatch bound to the host build of the device step body
(tests/emu/salp_emu.cu) instead of libsalp_b200.so, for the GPU-less container."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import build_emu  # noqa: E402

from grasp_lab_salp_b200 import _lib  # noqa: E402
from grasp_lab_salp_b200.batch import SalpBatch  # noqa: E402

_EMU_SYMBOLS = {"salp_create", "salp_destroy", "salp_num_envs", "salp_obs_dim", "salp_last_error",
                "salp_build_info", "salp_reset_host", "salp_step_host", "salp_set_scene_pool",
                "salp_get_state", "salp_set_state", "salp_check", "salp_launch_count", "salp_trace_cycle"}
_cdll = None


def emu_cdll():
    global _cdll
    if _cdll is None:
        _cdll = _lib.bind(C.CDLL(build_emu.build()), names=_EMU_SYMBOLS)
    return _cdll


def EmuBatch(num_envs, params=None, **kw):
    return SalpBatch(num_envs, params, _cdll=emu_cdll(), **kw)


_cdll_lane = None


def emu_lane_cdll():
    global _cdll_lane
    if _cdll_lane is None:
        _cdll_lane = _lib.bind(C.CDLL(build_emu.build(lane_only=True)), names=_EMU_SYMBOLS)
    return _cdll_lane
