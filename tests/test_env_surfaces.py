"""The reference-facing surfaces: gymnasium-style SalpRobotEnv and the SB3-VecEnv-shaped
SalpCudaVecEnv.  CPU variants run on the host build of the kernel body (tests/emu); the GPU
variants (marked gpu) run the same checks on the CUDA library."""
import numpy as np
import pytest

from grasp_lab_salp_b200 import PRECISION_F64, Nozzle, Robot, SalpCudaVecEnv, SalpRobotEnv
from oracle.salp_oracle import OracleVecEnv
from parity import golden_params, load_golden

REWARD_KEYS = ["rewards/track", "rewards/heading", "rewards/smooth", "rewards/yaw", "rewards/time",
               "rewards/sideslip", "rewards/obstacle"]


def _cdll(use_gpu):
    if use_gpu:
        return None
    from emu_backend import emu_cdll
    return emu_cdll()


def make_env(cdll):
    """Verbatim body of the reference's make_env() (src/train_robot.py:11-21) on the stand-in classes."""
    nozzle = Nozzle(length1=0.05, length2=0.05, length3=0.05, area=0.00016, mass=1.0)
    robot = Robot(dry_mass=1.0, init_length=0.3, init_width=0.15, max_contraction=0.06, nozzle=nozzle)
    robot.nozzle.set_angles(angle1=0.0, angle2=0.0)
    robot.set_environment(density=1000)
    env = SalpRobotEnv(render_mode=None, robot=robot, precision=PRECISION_F64, _cdll=cdll)
    return env


def _single_env_replays_reference_trace(cdll):
    g = load_golden("ref_fixed10.npz")
    env = make_env(cdll)
    assert env.action_space.shape == (3,) and env.observation_space.shape == (10,)
    env.set_scene(g["targets"][0, 0], g["obstacles"][0, 0])
    obs, info = env.reset(seed=0)
    assert info == {} and obs.dtype == np.float32
    np.testing.assert_allclose(obs, g["first_obs"][0], rtol=1e-6, atol=1e-7)
    np.testing.assert_array_equal(env.target_point, g["targets"][0, 0])
    for t in range(g["actions"].shape[1]):
        obs, rew, done, trunc, info = env.step(g["actions"][0, t])
        assert isinstance(rew, float) and isinstance(done, bool) and isinstance(trunc, bool)
        np.testing.assert_allclose(obs, g["obs"][0, t], rtol=2e-6, atol=1e-6)
        assert rew == pytest.approx(g["reward"][0, t], rel=1e-8, abs=1e-8)
        assert done == bool(g["terminated"][0, t]) and trunc == bool(g["truncated"][0, t])
        for j, k in enumerate(REWARD_KEYS):
            assert info[k] == pytest.approx(g["terms"][0, t, j], rel=1e-8, abs=1e-9)
        assert info["position_history"] == [] and env.robot.cycle == g["cycle"][0, t]
        np.testing.assert_allclose(env.robot.position_world[:2], g["state"][0, t, :2], rtol=1e-9, atol=1e-12)
        if done or trunc:
            env.reset()
    env.close()


def _vec_env_has_sb3_semantics(cdll, n=24, T=40):
    g = load_golden("ref_random.npz")
    params = golden_params(g, precision=PRECISION_F64)
    venv = SalpCudaVecEnv(n, params, seed=3, info_mode="full", _cdll=cdll)
    orc = OracleVecEnv(n, params, seed=3)
    obs = venv.reset()
    np.testing.assert_allclose(obs, orc.reset(), rtol=1e-6, atol=1e-7)
    assert obs.shape == (n, 10) and obs.dtype == np.float32 and len(venv.reset_infos) == n
    assert venv.env_is_wrapped(object) == [False] * n and venv.get_attr("render_mode") == [None] * n
    rng = np.random.default_rng(0)
    returns = np.zeros(n)
    lengths = np.zeros(n, int)
    episodes = 0
    for t in range(T):
        a = rng.uniform([0, 0, -1], [1, 1, 1], size=(n, 3)).astype(np.float32)
        venv.step_async(a)
        obs, rew, dones, infos = venv.step_wait()
        o_obs, o_rew, o_term, o_trunc = orc.step(a, auto_reset=True)
        assert rew.dtype == np.float32 and dones.dtype == bool and len(infos) == n
        np.testing.assert_array_equal(dones, (o_term | o_trunc).astype(bool))
        np.testing.assert_allclose(obs, o_obs, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(rew, o_rew.astype(np.float32), rtol=1e-5, atol=1e-4)
        returns += o_rew
        lengths += 1
        for i in range(n):
            info = infos[i]
            assert set(REWARD_KEYS) <= set(info)
            if dones[i]:
                episodes += 1
                np.testing.assert_allclose(info["terminal_observation"], orc.terminal_obs[i], rtol=1e-5, atol=1e-6)
                assert info["TimeLimit.truncated"] == bool(o_trunc[i] and not o_term[i])
                assert info["episode"]["l"] == lengths[i]
                assert info["episode"]["r"] == pytest.approx(returns[i], rel=1e-6, abs=1e-6)
                assert info["path_length"] == pytest.approx(orc.metrics[i, 2], rel=1e-6, abs=1e-9)
                assert info["final_distance"] == pytest.approx(orc.metrics[i, 5], rel=1e-6, abs=1e-9)
                returns[i] = 0.0
                lengths[i] = 0
            else:
                assert "terminal_observation" not in info and "episode" not in info
    assert episodes > 0
    # lazy infos: unfinished envs share one empty dict
    lazy = SalpCudaVecEnv(8, params, seed=3, info_mode="lazy", _cdll=cdll)
    lazy.reset()
    _, _, d, infos = lazy.step(np.tile(np.array([[0.5, 0.1, 0.2]], np.float32), (8, 1)))
    assert all((infos[i] == {}) for i in range(8) if not d[i])
    with pytest.raises(ValueError):
        lazy.step_async(np.zeros((3, 3), np.float32))
    venv.close()
    lazy.close()


def test_single_env_replays_reference_trace_emu():
    _single_env_replays_reference_trace(_cdll(False))


def test_vec_env_has_sb3_semantics_emu():
    _vec_env_has_sb3_semantics(_cdll(False))


@pytest.mark.gpu
def test_single_env_replays_reference_trace_gpu():
    _single_env_replays_reference_trace(_cdll(True))


@pytest.mark.gpu
def test_vec_env_has_sb3_semantics_gpu():
    _vec_env_has_sb3_semantics(_cdll(True), n=256, T=40)


def _history_feed(cdll):
    """record=True: info carries the per-substep histories of the cycle (robot.py:681-776); the
    last row is the state the step leaves behind, the row count is K, and tracing does not
    advance the env."""
    g = load_golden("ref_fixed10.npz")
    env = make_env(cdll)
    env.enable_history_recording()
    env.set_scene(g["targets"][0, 0], g["obstacles"][0, 0])
    env.reset()
    for t in range(4):
        a = g["actions"][0, t]
        before = env.robot.position_world.copy()
        h = env._batch.trace_cycle(0, a)
        np.testing.assert_array_equal(env.robot.position_world, before)          # not advanced
        obs, rew, done, trunc, info = env.step(a)
        K = int(g["K"][0, t])
        assert h["substeps"] == K and info["position_history"].shape == (K, 3)
        assert info["length_history"].shape == (K,) and info["width_history"].shape == (K,)
        np.testing.assert_allclose(info["position_history"][-1], env.robot.position_world, rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(info["position_history"][-1][:2], g["state"][0, t, :2], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(env.last_history["euler_angle"][-1], env.robot.euler_angle, rtol=1e-12, atol=1e-15)
        assert info["length_history"].min() >= 0.3 - 0.06 * a[0] - 1e-6 and abs(info["length_history"][-1] - 0.3) < 1e-12
        np.testing.assert_allclose(info["length_history"] + info["width_history"], 0.45, rtol=1e-12)
    env.close()


def test_history_feed_emu():
    _history_feed(_cdll(False))


@pytest.mark.gpu
def test_history_feed_gpu():
    _history_feed(_cdll(True))
