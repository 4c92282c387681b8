"""GPU (B200): the CUDA path, called through the C ABI, against the golden traces of the live
Python reference, against the C oracle in lockstep, and through size-independent properties
at BASELINE.json's full size (4096 envs)."""
import numpy as np
import pytest

from grasp_lab_salp_b200 import PRECISION_F64, PRECISION_MIXED, SalpBatch
from oracle.salp_oracle import OracleVecEnv
from parity import (check_blowup_golden, TOL_F64, MIXED_FLOORS, TOL_MIXED, TOL_MIXED_FREE_RUN, golden_params, load_golden, lockstep_compare,
                    replay_golden, sample_scene_pool)

pytestmark = pytest.mark.gpu
GOLDENS = ["ref_fixed10.npz", "ref_edge.npz", "ref_random.npz", "ref_clipped.npz"]


def uniform_actions(rng, T, n):
    return rng.uniform([0, 0, -1], [1, 1, 1], size=(T, n, 3)).astype(np.float32)


def clipped_actions(rng, T, n):
    return np.clip(rng.normal(size=(T, n, 3)), [0, 0, -1], [1, 1, 1]).astype(np.float32)


@pytest.mark.parametrize("name", GOLDENS)
def test_f64_kernel_matches_reference_trace(name):
    g = load_golden(name)
    env = SalpBatch(g["actions"].shape[0], golden_params(g, precision=PRECISION_F64))
    report = {}
    replay_golden(env, g, report=report, **TOL_F64)
    env.check()
    print(name, report)


@pytest.mark.parametrize("name", GOLDENS)
def test_mixed_kernel_matches_reference_trace(name):
    g = load_golden(name)
    env = SalpBatch(g["actions"].shape[0], golden_params(g, precision=PRECISION_MIXED))
    report = {}
    replay_golden(env, g, report=report, **TOL_MIXED_FREE_RUN)
    env.check()
    print(name, report)


def _pair(n, precision, seed=3, P=6, pool=True, threads=8):
    g = load_golden("ref_random.npz")
    params = golden_params(g, precision=precision)
    prod = SalpBatch(n, params, seed=seed)
    orc = OracleVecEnv(n, params, seed=seed, threads=threads)
    if pool:
        t, o = sample_scene_pool(np.random.default_rng(seed), n, P)
        prod.set_scene_pool(t, o)
        orc.set_scene_pool(t, o)
    return prod, orc


@pytest.mark.parametrize("kind", ["uniform", "clipped"])
def test_mixed_per_step_tolerance_vs_oracle(kind):
    """North-star tolerance: ONE env-step from identical state, fp32 CUDA kernel vs the float64
    oracle.  K, cycle, phase, done/truncated, reset indices bit-exact; 1e-5 relative (vs
    max(|ref|, 0.1)) on position, velocity, heading, body shape, obs; reward vs max(|r|, 10)."""
    n, T = 512, 16
    prod, orc = _pair(n, PRECISION_MIXED)
    rng = np.random.default_rng(11)
    acts = uniform_actions(rng, T, n) if kind == "uniform" else clipped_actions(rng, T, n)
    report = {}
    lockstep_compare(prod, orc, acts, resync=True, rtol=TOL_MIXED["rtol"], floor=TOL_MIXED["floor"], floors=MIXED_FLOORS, report=report)
    prod.check()
    print(kind, report)


def test_f64_free_running_with_autoreset_and_philox_scenes_vs_oracle():
    """No scene pool: the built-in Philox sampler of the kernel and of the oracle must produce
    the same float32 targets/obstacles, so 40 free-running steps with in-kernel auto-reset stay
    in lockstep (same reset indices, every counter identical)."""
    n, T = 256, 40
    prod, orc = _pair(n, PRECISION_F64, pool=False)
    acts = uniform_actions(np.random.default_rng(5), T, n)
    hist = lockstep_compare(prod, orc, acts, resync=False, rtol=1e-9, floor=1e-3, obs_floor=1e-3, reward_floor=1.0)
    worst = {k: max(h[k] for h in hist) for k in hist[0]}
    print(worst)
    assert max(worst.values()) < 1e-7, worst
    assert orc.get_state("episode_index").max() > 1
    prod.check()


def test_sort_by_k_is_bit_identical():
    n, T = 2048, 6
    g = load_golden("ref_random.npz")
    acts = uniform_actions(np.random.default_rng(2), T, n)
    outs = []
    for sort in (False, True):
        env = SalpBatch(n, golden_params(g, precision=PRECISION_MIXED), seed=9)
        env.reset()
        rec = []
        for t in range(T):
            obs, rew, term, trunc = env.step(acts[t], auto_reset=True, sort_by_k=sort, pipeline=False)
            rec.append((obs.copy(), rew.copy(), term.copy(), trunc.copy(), env.substeps.copy(),
                        env.get_state("posw_x"), env.get_state("euler_z")))
        env.check()
        outs.append(rec)
    for a, b in zip(*outs):
        for x, y in zip(a, b):
            np.testing.assert_array_equal(x, y)


def test_device_face_equals_host_face():
    import torch
    n, T = 1024, 5
    g = load_golden("ref_random.npz")
    acts = uniform_actions(np.random.default_rng(4), T, n)
    host = SalpBatch(n, golden_params(g), seed=1)
    dev = SalpBatch(n, golden_params(g), seed=1)
    host.reset()
    obs_d = dev.reset_device()
    np.testing.assert_array_equal(obs_d.cpu().numpy(), host.obs)
    for t in range(T):
        o, r, te, tr = host.step(acts[t], auto_reset=True)
        od, rd, ted, trd = dev.step_device(torch.from_numpy(acts[t]).cuda(), auto_reset=True, extras=True)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(od.cpu().numpy(), o)
        np.testing.assert_array_equal(rd.cpu().numpy(), r)
        np.testing.assert_array_equal(ted.cpu().numpy(), te)
        np.testing.assert_array_equal(trd.cpu().numpy(), tr)
        np.testing.assert_array_equal(dev.dev["terminal_obs"].cpu().numpy(), host.terminal_obs)
        np.testing.assert_array_equal(dev.dev["substeps"].cpu().numpy(), host.substeps)
    v = dev.state_tensor("posw_x")
    np.testing.assert_array_equal(v.cpu().numpy(), host.get_state("posw_x"))


@pytest.mark.parametrize("sort", [False, True])
def test_zero_copy_host_path_equals_staged_host_path(sort):
    """salp_step_host writes straight into page-locked caller buffers (mapped host memory, rows
    assembled in shared memory) and stages through device buffers for pageable ones: same bits.
    The last warp is ragged (n % 32 != 0) and the batch ends/auto-resets episodes on the way."""
    n, T = 1000, 12
    g = load_golden("ref_random.npz")
    acts = uniform_actions(np.random.default_rng(8), T, n)
    pinned = SalpBatch(n, golden_params(g), seed=3)
    paged = SalpBatch(n, golden_params(g), seed=3)
    for name in ("obs", "terminal_obs", "reward", "terminated", "truncated"):
        setattr(paged, name, np.full_like(getattr(paged, name), 7))       # plain numpy: not page-locked
    a_pinned = pinned.host_buffer((n, 3), np.float32)
    np.testing.assert_array_equal(pinned.reset(), paged.reset())
    ended = 0
    for t in range(T):
        a_pinned[:] = acts[t]
        o, r, te, tr = pinned.step(a_pinned, auto_reset=True, sort_by_k=sort)
        o2, r2, te2, tr2 = paged.step(acts[t].copy(), auto_reset=True, sort_by_k=sort)
        for x, y in ((o, o2), (r, r2), (te, te2), (tr, tr2), (pinned.terminal_obs, paged.terminal_obs),
                     (pinned.substeps, paged.substeps), (pinned.terms, paged.terms)):
            np.testing.assert_array_equal(x, y)
        ended += int((te | tr).sum())
    assert ended > 0
    pinned.check()
    paged.check()


def test_full_size_mixed_vs_f64_trajectory_equivalence():
    """BASELINE config 2 shape (4096 envs, random actions, auto-reset), free-running: the fp32
    kernel against the float64 kernel on every env.  Integer quantities must agree exactly on
    (almost) every env-step -- a flag can only differ where a float64 threshold (target radius,
    5 m bound, obstacle contact) is crossed within fp32 noise; such an env is re-synchronised
    and counted.  compare_trajectories.py metrics (position / velocity L2, |yaw| error;
    reference src/compare_trajectories.py:77-86) are reported over the horizon."""
    n, T = 4096, 60
    g = load_golden("ref_random.npz")
    mixed = SalpBatch(n, golden_params(g, precision=PRECISION_MIXED), seed=7)
    f64 = SalpBatch(n, golden_params(g, precision=PRECISION_F64), seed=7)
    acts = uniform_actions(np.random.default_rng(8), T, n)
    mixed.reset()
    f64.reset()
    flag_mismatch = unstable_cycles = 0
    pos_err, vel_err, yaw_err = [], [], []
    for t in range(T):
        om, rm, tem, trm = mixed.step(acts[t], auto_reset=True)
        of, rf, tef, trf = f64.step(acts[t], auto_reset=True)
        np.testing.assert_array_equal(mixed.substeps, f64.substeps)        # K is decided in fp64/fp32-exact code
        bad = (tem != tef) | (trm != trf)
        flag_mismatch += int(bad.sum())
        # a cycle next to the integrator's stability limit (a0 ~ 0.09) amplifies any
        # rounding difference by many orders of magnitude: such an env is counted and re-synchronised
        diverged = ~bad & ~(np.abs(rm - rf) <= 2e-3 + 1e-4 * np.abs(rf))
        unstable_cycles += int(diverged.sum())
        bad = bad | diverged
        ok = ~bad
        px = mixed.get_state("posw_x") - f64.get_state("posw_x")
        py = mixed.get_state("posw_y") - f64.get_state("posw_y")
        vx = mixed.get_state("vel_x") - f64.get_state("vel_x")
        vy = mixed.get_state("vel_y") - f64.get_state("vel_y")
        yaw = mixed.get_state("euler_z") - f64.get_state("euler_z")
        pos_err.append(np.hypot(px, py)[ok].max())
        vref = np.hypot(f64.get_state("vel_x"), f64.get_state("vel_y"))
        vel_err.append((np.hypot(vx, vy) / np.maximum(vref, 0.1))[ok].max())   # relative: blow-up-adjacent cycles spike |v|
        yaw_err.append(np.abs(yaw)[ok].max())
        if bad.any():      # re-synchronise the diverged envs from the float64 run
            from parity import all_columns
            for col in all_columns():
                if col.startswith("obstacle") and int(col[8]) >= 2:
                    continue
                v = mixed.get_state(col)
                v[bad] = f64.get_state(col)[bad]
                mixed.set_state(col, v)
    print(f"4096x{T}: flag mismatches {flag_mismatch}, diverged near the stability limit {unstable_cycles}, max pos err {max(pos_err):.2e} m, "
          f"max rel vel err {max(vel_err):.2e}, max yaw err {max(yaw_err):.2e} rad")
    assert flag_mismatch <= 2 and unstable_cycles <= 12      # of 245 760 env-steps
    assert max(pos_err) < 5e-5 and max(vel_err) < 5e-5 and max(yaw_err) < 5e-5     # free-running drift over 60 env-steps
    mixed.check()
    f64.check()


@pytest.mark.parametrize("n", [512, 8192, 12288])   # pipeline kernel everywhere / fused whole vs pipeline shards / fused everywhere
def test_shard_invariance(n):
    """Multi-GPU sharding rule: envs [0,N) on one handle == two handles of N/2 with
    env_id_offset, bit for bit (per-env Philox streams are keyed by the GLOBAL env id; results
    do not depend on which envs share a warp)."""
    T = 20 if n <= 4736 else 8
    g = load_golden("ref_random.npz")
    acts = uniform_actions(np.random.default_rng(6), T, n)
    whole = SalpBatch(n, golden_params(g), seed=21)
    lo = SalpBatch(n // 2, golden_params(g), seed=21, env_id_offset=0)
    hi = SalpBatch(n // 2, golden_params(g), seed=21, env_id_offset=n // 2)
    np.testing.assert_array_equal(whole.reset(), np.concatenate([lo.reset(), hi.reset()]))
    for t in range(T):
        o, r, te, tr = whole.step(acts[t], auto_reset=True)
        o1, r1, te1, tr1 = lo.step(acts[t, : n // 2], auto_reset=True)
        o2, r2, te2, tr2 = hi.step(acts[t, n // 2:], auto_reset=True)
        np.testing.assert_array_equal(o, np.concatenate([o1, o2]))
        np.testing.assert_array_equal(r, np.concatenate([r1, r2]))
        np.testing.assert_array_equal(te, np.concatenate([te1, te2]))
        np.testing.assert_array_equal(tr, np.concatenate([tr1, tr2]))


def test_non_finite_action_raises_range_error_not_a_hang():
    from grasp_lab_salp_b200 import SalpError
    env = SalpBatch(64, seed=0)
    env.reset()
    a = np.zeros((64, 3), np.float32)
    a[3, 1] = np.inf          # infinite coast time
    env.step(a)
    with pytest.raises(SalpError):
        env.check()


@pytest.mark.parametrize("precision", [PRECISION_F64, PRECISION_MIXED])
def test_kernel_cuts_exactly_the_episodes_where_the_reference_raises(precision):
    worst = check_blowup_golden(lambda n, g: SalpBatch(n, golden_params(g, precision=precision)))
    print("blow-up neighbourhood: worst final-pose error of the surviving cycles", worst)
    # these cycles sit next to the integrator's stability limit (transient |v| of 1e3 m/s and more):
    # fp32 rounding is amplified accordingly, hence 1e-3 here instead of the 1e-5 of regular cycles
    assert worst < (1e-9 if precision == PRECISION_F64 else 1e-3), worst


def test_pipeline_kernel_matches_fused_kernel():
    """Small batches run the three-warp producer/consumer kernel (salp_pipe_kernel.cuh): the same
    functions as the fused kernel split over three warps.  Every operation of the substep loop is
    explicitly rounded (fmaf / __fmul_rn / ...), so the two kernels agree bit for bit -- outputs
    and every state column, free-running over 12 steps with ragged blocks, K = 0 warps, blow-up
    cuts and auto-resets."""
    from grasp_lab_salp_b200.params import FIELDS
    n, T = 1000, 12           # not a multiple of 32: ragged last block
    g = load_golden("ref_random.npz")
    acts = uniform_actions(np.random.default_rng(12), T, n)
    acts[0, :8] = [[0, 0, 0], [1, 1, 1], [1, 0, -1], [0.088, 0, 0.5], [0.5, 0.5, 1e-4], [0.09, 0, 0], [0, 1, 1], [1, 1, 0]]
    acts[3, 32:64] = 0.0      # a whole warp with K = 0
    pipe = SalpBatch(n, golden_params(g), seed=2)
    fused = SalpBatch(n, golden_params(g), seed=2)
    np.testing.assert_array_equal(pipe.reset(), fused.reset())
    ended = 0
    for t in range(T):
        o1, r1, te1, tr1 = pipe.step(acts[t], auto_reset=True, pipeline=True)
        o2, r2, te2, tr2 = fused.step(acts[t], auto_reset=True, pipeline=False)
        for x, y in ((o1, o2), (r1, r2), (te1, te2), (tr1, tr2), (pipe.substeps, fused.substeps),
                     (pipe.terminal_obs, fused.terminal_obs), (pipe.terms, fused.terms)):
            np.testing.assert_array_equal(x, y)
        ended += int((te1 | tr1).sum())
    for col in FIELDS:
        np.testing.assert_array_equal(pipe.get_state(col), fused.get_state(col), err_msg=col)
    assert ended > 0
    pipe.check()
    fused.check()


@pytest.mark.parametrize("n,pipeline", [(1000, True), (1000, False), (24000, None)])
def test_axisymmetric_form_equals_general_form(n, pipeline):
    """SalpDerived.axisym: the form of the loop for axisymmetric coefficient sets (the defaults) vs
    the general form (SALP_STEP_GENERIC): bit-identical, in the pipeline kernel, the fused kernel
    and the fused K-sorted kernel."""
    from grasp_lab_salp_b200.params import FIELDS
    g = load_golden("ref_random.npz")
    T = 8 if n <= 4736 else 4
    acts = uniform_actions(np.random.default_rng(15), T, n)
    a, b = SalpBatch(n, golden_params(g), seed=6), SalpBatch(n, golden_params(g), seed=6)
    np.testing.assert_array_equal(a.reset(), b.reset())
    sort = n > 18944
    for t in range(T):
        ra = a.step(acts[t], auto_reset=True, pipeline=pipeline, sort_by_k=sort)
        rb = b.step(acts[t], auto_reset=True, pipeline=pipeline, sort_by_k=sort, generic=True)
        for x, y in zip(ra, rb):
            np.testing.assert_array_equal(x, y)
        np.testing.assert_array_equal(a.terms, b.terms)
    for col in FIELDS:
        np.testing.assert_array_equal(a.get_state(col), b.get_state(col), err_msg=col)
    a.check()
    b.check()


@pytest.mark.parametrize("n", [6001, 9472])
def test_pipeline_kernel_two_blocks_per_sm_matches_fused_kernel(n):
    """4737-9472 envs: two pipeline blocks share an SM and the second one rotates its warp roles
    (hardware warp slot + per-SM ticket, salp_pipe4_kernel.cuh).  Whatever roles the warps draw, the
    result is the fused kernel's, bit for bit; the hand-off tags are checked on the way."""
    from grasp_lab_salp_b200.params import FIELDS
    T = 6
    g = load_golden("ref_random.npz")
    acts = uniform_actions(np.random.default_rng(21), T, n)
    acts[2, 64:128] = 0.0
    pipe, fused = SalpBatch(n, golden_params(g), seed=4), SalpBatch(n, golden_params(g), seed=4)
    np.testing.assert_array_equal(pipe.reset(), fused.reset())
    for t in range(T):
        r1 = pipe.step(acts[t], auto_reset=True, check_handoff=(t % 2 == 0))
        assert pipe.last_step_kernel == "salp_step_kernel_pipe4"
        r2 = fused.step(acts[t], auto_reset=True, pipeline=False)
        assert fused.last_step_kernel != "salp_step_kernel_pipe4"
        for x, y in zip(r1, r2):
            np.testing.assert_array_equal(x, y)
        np.testing.assert_array_equal(pipe.substeps, fused.substeps)
    pipe.check()
    for col in FIELDS:
        np.testing.assert_array_equal(pipe.get_state(col), fused.get_state(col), err_msg=col)


def test_pipeline_kernel_is_deterministic_at_full_size():
    """Two handles, same seed, same actions, BASELINE's 4096 envs (every SM busy with one
    three-warp block), 40 free-running steps: bit-identical outputs and state.  A missed hand-off
    between the producer and consumer warps (a shared-memory race) would show up here as a
    run-to-run difference."""
    from grasp_lab_salp_b200.params import FIELDS
    n, T = 4096, 40
    g = load_golden("ref_random.npz")
    acts = uniform_actions(np.random.default_rng(77), T, n)
    a, b = SalpBatch(n, golden_params(g), seed=5), SalpBatch(n, golden_params(g), seed=5)
    np.testing.assert_array_equal(a.reset(), b.reset())
    for t in range(T):
        ra = a.step(acts[t], auto_reset=True, pipeline=True)
        rb = b.step(acts[t], auto_reset=True, pipeline=True)
        for x, y in zip(ra, rb):
            np.testing.assert_array_equal(x, y)
        np.testing.assert_array_equal(a.substeps, b.substeps)
    for col in FIELDS:
        np.testing.assert_array_equal(a.get_state(col), b.get_state(col), err_msg=col)
    a.check()
    b.check()


def test_pipeline_kernel_short_cycles_and_chunk_boundaries():
    """Cycles of K = 0, 1, 2, ... substeps (contraction 0, coast chosen so that total = K * dt) put
    the end of the cycle on every position of the pipeline kernel's 8-substep hand-off chunks and
    of the 32-substep flush boundary; the same coasts again behind a long, a tiny and no
    contraction with a nozzle turn.  Pipeline kernel (the default at this size) vs the float64
    oracle at the per-step tolerance, and vs the fused kernel."""
    ks = np.concatenate([np.arange(0, 42), [63, 64, 65, 66, 95, 96, 97, 127, 128, 129]])
    m = len(ks)
    n = 4 * m
    acts = np.zeros((n, 3), np.float32)
    # refill(0) = -0.45, jet(0) = -0.125: total = 0 + jet + coast  ->  coast = 0.125 + (K - 0.5) * dt
    coast = 0.125 + (np.tile(ks, 4) - 0.5) * 0.01
    acts[:, 1] = np.where(np.tile(ks, 4) > 0, coast / 10.0, 0.0)
    acts[m:2 * m, 0] = 0.3            # a real contraction (long shape motion) in front of the same coasts
    acts[2 * m:3 * m, 0] = 0.05       # a tiny one
    acts[3 * m:, 2] = 0.7             # nozzle turn only
    prod, orc = _pair(n, PRECISION_MIXED)
    probe = OracleVecEnv(n, prod.params, seed=3)
    probe.reset()
    probe.step(acts)
    np.testing.assert_array_equal(probe.substeps[:m], ks)          # the crafted coasts give K = 0, 1, 2, ...
    report = {}
    lockstep_compare(prod, orc, np.repeat(acts[None], 3, axis=0), resync=True, rtol=TOL_MIXED["rtol"],
                     floor=TOL_MIXED["floor"], floors=MIXED_FLOORS, report=report)
    prod.check()
    print(report)
    pipe, fused = SalpBatch(n, prod.params, seed=4), SalpBatch(n, prod.params, seed=4)
    pipe.reset(), fused.reset()
    for t in range(2):
        o1, r1, te1, tr1 = pipe.step(acts, auto_reset=True, pipeline=True)
        o2, r2, te2, tr2 = fused.step(acts, auto_reset=True, pipeline=False)
        np.testing.assert_array_equal(pipe.substeps, fused.substeps)
        np.testing.assert_array_equal(te1, te2)
        np.testing.assert_array_equal(tr1, tr2)
        np.testing.assert_allclose(o1, o2, rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(r1, r2, rtol=1e-5, atol=2e-4)


def test_other_obstacle_count_and_masked_device_reset():
    import torch
    g = load_golden("ref_random.npz")
    n, T = 256, 6
    params = golden_params(g, precision=PRECISION_MIXED, num_obstacles=5)
    prod = SalpBatch(n, params, seed=7)
    orc = OracleVecEnv(n, params, seed=7, threads=1)
    assert prod.obs_dim == 16
    acts = uniform_actions(np.random.default_rng(3), T, n)
    lockstep_compare(prod, orc, acts, resync=True, rtol=TOL_MIXED["rtol"], floor=TOL_MIXED["floor"], floors=MIXED_FLOORS, num_obstacles=5)
    mask = np.zeros(n, np.uint8)
    mask[::7] = 1
    obs_o = orc.reset(mask).copy()
    prod.dev["obs"].copy_(torch.from_numpy(prod.obs).cuda())
    obs_d = prod.reset_device(torch.from_numpy(mask).cuda())
    torch.cuda.synchronize()
    np.testing.assert_allclose(obs_d.cpu().numpy()[mask == 1], obs_o[mask == 1], rtol=1e-6, atol=1e-6)
    np.testing.assert_array_equal(prod.get_state("episode_index"), orc.get_state("episode_index"))
    np.testing.assert_array_equal(prod.get_state("cycle"), orc.get_state("cycle"))


def test_pipeline_handoff_tags_at_full_size():
    """SALP_STEP_CHECK_HANDOFF: every shared-memory ring row of the warp-specialised kernel carries the
    substep index it was produced for and each consuming warp verifies it (salp_pipe4_kernel.cuh).
    BASELINE's 4096 envs (one block on almost every SM, all four warps busy), 60 free-running steps
    with auto-reset, plus ragged / K = 0 / short-cycle batches: no wrong tag, and the checked kernel
    gives the same bits as the unchecked one.  (compute-sanitizer's racecheck is closed on the
    measurement pool: `compute-sanitizer is closed on this pool and stays closed`.)"""
    from grasp_lab_salp_b200.params import FIELDS
    g = load_golden("ref_random.npz")
    for n, T in ((4096, 60), (1000, 12), (77, 12)):
        acts = uniform_actions(np.random.default_rng(31), T, n)
        acts[1, : min(n, 64)] = 0.0                       # K = 0 warps
        acts[2, :, 1] *= 0.02                             # short cycles: the shape moves almost all the time
        a, b = SalpBatch(n, golden_params(g), seed=5), SalpBatch(n, golden_params(g), seed=5)
        np.testing.assert_array_equal(a.reset(), b.reset())
        for t in range(T):
            ra = a.step(acts[t], auto_reset=True, check_handoff=True)
            rb = b.step(acts[t], auto_reset=True)
            assert a.last_step_kernel == "salp_step_kernel_pipe4"
            for x, y in zip(ra, rb):
                np.testing.assert_array_equal(x, y)
        a.check()                                         # raises SALP_ERR_HANDOFF on a wrong tag
        for col in FIELDS:
            np.testing.assert_array_equal(a.get_state(col), b.get_state(col), err_msg=col)
        a.close()
        b.close()


def test_chunked_host_step_equals_single_launch():
    """salp_step_host, K-sorted, large batch with page-locked buffers: the batch is stepped in
    contiguous env ranges whose result copies overlap the next range's kernel (salp_capi.cu,
    step_host_ranges).  Same bits as the one-launch staged path (pageable buffers), including the
    ragged last range, the built-in Philox scenes (global env ids) and auto-reset."""
    n, T = 300_003, 3           # 2 ranges of 150 016 envs, the last one ragged
    g = load_golden("ref_random.npz")
    acts = uniform_actions(np.random.default_rng(18), T, n)
    pinned = SalpBatch(n, golden_params(g), seed=3)
    paged = SalpBatch(n, golden_params(g), seed=3)
    for name in ("obs", "terminal_obs", "reward", "terminated", "truncated"):
        setattr(paged, name, np.full_like(getattr(paged, name), 7))       # plain numpy: not page-locked
    a_pinned = pinned.host_buffer((n, 3), np.float32)
    np.testing.assert_array_equal(pinned.reset(), paged.reset())
    for t in range(T):
        a_pinned[:] = acts[t]
        o, r, te, tr = pinned.step(a_pinned, auto_reset=True, sort_by_k=True)
        o2, r2, te2, tr2 = paged.step(acts[t].copy(), auto_reset=True, sort_by_k=True)
        for x, y in ((o, o2), (r, r2), (te, te2), (tr, tr2), (pinned.terminal_obs, paged.terminal_obs),
                     (pinned.substeps, paged.substeps), (pinned.terms, paged.terms)):
            np.testing.assert_array_equal(x, y)
    np.testing.assert_array_equal(pinned.get_state("episode_index"), paged.get_state("episode_index"))
    np.testing.assert_array_equal(pinned.get_state("posw_x"), paged.get_state("posw_x"))
    pinned.check()
    paged.check()
