"""Default-off robustness switches of the reference: action / observation
randomisation, per-cycle dynamics randomisation, OU force/torque disturbances, latency.  They
draw from the global np.random in the reference, so parity is STATISTICAL: the distribution of
the outcome of 3 fixed cycles over 320 seeds of the live reference (tests/golden/ref_randstats.npz,
tools/gen_golden.py randstats) against 4096 envs of the kernel body."""
import numpy as np
import pytest

from grasp_lab_salp_b200 import PRECISION_F64, PRECISION_MIXED, SalpBatch, SalpError
from grasp_lab_salp_b200.params import (RAND_ACTION, RAND_DISTURBANCE, RAND_DYNAMICS, RAND_LATENCY,
                                        RAND_OBSERVATION)
from parity import golden_params, load_golden

MODES = {"dynamics": RAND_DYNAMICS, "disturbance": RAND_DISTURBANCE, "action": RAND_ACTION,
         "observation": RAND_OBSERVATION}


def _run(flag, n, cdll, seed=0):
    g = load_golden("ref_randstats.npz")
    p = golden_params(g, precision=PRECISION_MIXED)
    p.randomization = flag
    env = SalpBatch(n, p, seed=seed, _cdll=cdll)
    env.set_scene_pool(np.tile(np.array([[[1.8, 1.2]]], np.float32), (n, 1, 1)),
                       np.tile(np.array([[[[-1.5, -1.0], [1.5, -1.0]]]], np.float32), (n, 1, 1, 1)))
    env.reset()
    obs = None
    for a in g["actions"]:
        obs, _, _, _ = env.step(np.tile(a[None], (n, 1)))
    cols = [env.get_state(c) for c in ("posw_x", "posw_y", "euler_z", "vel_x", "vel_y", "angvel_z")]
    out = np.stack(cols + [obs[:, k].astype(np.float64) for k in range(6)] + [env.get_state("nozzle_yaw").astype(np.float64)], 1)
    cyc = env.get_state("cycle")
    env.close()
    return out, cyc, g


def _check(mode, cdll, n):
    got, cyc, g = _run(MODES[mode], n, cdll)
    ref = g[f"samples_{mode}"]
    base = g["samples_none"][0]
    m = ref.shape[0]
    assert (cyc == 3).all()
    # which outcome columns does this switch move?
    moved = list(range(13)) if mode != "observation" else list(range(6, 12))
    if mode in ("dynamics", "disturbance"):
        moved = list(range(12))           # the commanded nozzle yaw is untouched
    for j in range(13):
        rs, gs = ref[:, j].std(), got[:, j].std()
        if j not in moved:
            np.testing.assert_allclose(got[:, j], base[j], rtol=2e-5, atol=2e-6, err_msg=f"{mode} col {j} should not move")
            continue
        if rs < 1e-9:     # the reference's clip quirk: a negative value is "randomised" to exactly v (1 + u)
            np.testing.assert_allclose(got[:, j], ref[0, j], rtol=2e-5, atol=2e-6, err_msg=f"{mode} col {j}")
            continue
        assert gs > 0, (mode, j)
        # mean: within 4 standard errors of the reference sample mean (+ a sliver for fp32)
        tol = 4.0 * rs / np.sqrt(m) + 4.0 * gs / np.sqrt(n) + 1e-6
        assert abs(got[:, j].mean() - ref[:, j].mean()) < tol, (mode, j, got[:, j].mean(), ref[:, j].mean(), tol)
        # spread: same standard deviation within the sampling error of 320 samples (~ +-12 % at 3 sigma)
        assert 0.8 < gs / rs < 1.25, (mode, j, gs, rs)


@pytest.mark.parametrize("mode", list(MODES))
def test_randomisation_statistics_match_reference_emu(mode):
    from emu_backend import emu_cdll
    _check(mode, emu_cdll(), 1024)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", list(MODES))
def test_randomisation_statistics_match_reference_gpu(mode):
    _check(mode, None, 8192)


def test_switches_off_is_the_deterministic_path_and_latency_bumps_cycle():
    from emu_backend import emu_cdll
    got0, cyc0, g = _run(0, 8, emu_cdll())
    np.testing.assert_allclose(got0, np.tile(g["samples_none"][0], (8, 1)), rtol=2e-5, atol=2e-6)
    _, cyc, _ = _run(RAND_LATENCY, 8, emu_cdll())
    assert (cyc == 6).all()                      # set_control runs twice per env-step (salp_robot_env.py:294-297)
    got_a, _, _ = _run(RAND_ACTION, 64, emu_cdll(), seed=1)
    got_b, _, _ = _run(RAND_ACTION, 64, emu_cdll(), seed=1)
    got_c, _, _ = _run(RAND_ACTION, 64, emu_cdll(), seed=2)
    np.testing.assert_array_equal(got_a, got_b)          # counter-based streams: reproducible
    assert not np.array_equal(got_a, got_c)
    # commanded yaw: U(0.9, 1.1) x (0.2 * pi/2) for the last action
    ratio = got_a[:, 12] / (0.2 * np.pi / 2)
    assert ratio.min() > 0.899 and ratio.max() < 1.101 and ratio.std() > 0.04


@pytest.mark.gpu
def test_randomisation_needs_the_mixed_kernel():
    g = load_golden("ref_randstats.npz")
    p = golden_params(g, precision=PRECISION_F64)
    p.randomization = RAND_DYNAMICS
    with pytest.raises(SalpError):
        SalpBatch(4, p)
