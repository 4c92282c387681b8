"""CPU, world_size 2 over gloo: the multi-GPU sharding rule.  Each rank steps its contiguous
shard (host build of the kernel body, tests/emu) with env_id_offset = shard offset and NO
data-path collective; rank 0 gathers and checks that the union equals the single-process run
bit for bit (the built-in Philox scene streams are keyed by the global env id)."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from grasp_lab_salp_b200.distributed import shard_for_rank

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOTAL, T = 37, 8          # odd on purpose: ragged shards


def test_shard_for_rank_covers_everything_once():
    for total in (1, 7, 37, 4096):
        for world in (1, 2, 3, 8):
            spans = [shard_for_rank(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (o1, c1), (o2, _) in zip(spans, spans[1:]):
                assert o1 + c1 == o2
    with pytest.raises(ValueError):
        shard_for_rank(8, 2, 2)


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch
    from emu_backend import emu_cdll
    from grasp_lab_salp_b200.distributed import allreduce_scalar, make_shard, world_info
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert world_info() == (rank, rank, world)
    shard = make_shard(TOTAL, seed=5, device=0, _cdll=emu_cdll())
    offset, count = shard_for_rank(TOTAL, rank, world)
    assert shard.num_envs == count
    acts = np.random.default_rng(1).uniform([0, 0, -1], [1, 1, 1], size=(T, TOTAL, 3)).astype(np.float32)
    obs = [shard.reset().copy()]
    steps = 0
    for t in range(T):
        o, r, te, tr = shard.step(acts[t, offset:offset + count], auto_reset=True)
        obs.append(np.concatenate([o, r[:, None], te[:, None].astype(np.float32), tr[:, None].astype(np.float32)], 1))
        steps += count
    total_steps = allreduce_scalar(float(steps), "sum")
    slowest = allreduce_scalar(float(rank + 1), "max")
    assert total_steps == TOTAL * T and slowest == world
    gathered = [None] * world
    dist.gather_object([x.tolist() for x in obs], gathered if rank == 0 else None, dst=0)
    if rank == 0:
        out.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_equal_the_single_process_run():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    gathered = out.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from emu_backend import EmuBatch
    whole = EmuBatch(TOTAL, seed=5)
    acts = np.random.default_rng(1).uniform([0, 0, -1], [1, 1, 1], size=(T, TOTAL, 3)).astype(np.float32)
    ref = [whole.reset().copy()]
    for t in range(T):
        o, r, te, tr = whole.step(acts[t], auto_reset=True)
        ref.append(np.concatenate([o, r[:, None], te[:, None].astype(np.float32), tr[:, None].astype(np.float32)], 1))
    for t in range(T + 1):
        got = np.concatenate([np.array(gathered[0][t], np.float32), np.array(gathered[1][t], np.float32)])
        np.testing.assert_array_equal(got, ref[t].astype(np.float32))


def _ppo_worker(rank, world, port, out, recurrent=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch
    from emu_backend import emu_cdll
    from grasp_lab_salp_b200.distributed import make_shard
    from grasp_lab_salp_b200.ppo import PPO, HostEnv, PPOConfig, RecurrentPPO
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shard = make_shard(13, seed=2, device=0, _cdll=emu_cdll())            # 7 + 6 envs: uneven shards
    if recurrent:      # minibatches are env sequences: 2 of them per step; BPTT through lstm_seq.LstmSequence
        algo = RecurrentPPO(HostEnv(shard), PPOConfig(n_steps=4, batch_size=8, n_epochs=2, seed=3))
    else:
        algo = PPO(HostEnv(shard), PPOConfig(n_steps=4, batch_size=8, n_epochs=2, seed=3))
    noise = torch.randn(3, generator=algo.gen)                             # per-rank exploration noise
    stats = algo.learn(2 * 4 * 13)
    w = torch.cat([p.detach().reshape(-1) for p in algo.policy.parameters()])
    gathered = [None] * world
    dist.gather_object((w.tolist(), noise.tolist(), algo._upd_calls, stats.env_steps), gathered if rank == 0 else None, dst=0)
    if rank == 0:
        out.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_ppo_with_uneven_shards_keeps_ranks_in_step():
    """PPO over gloo, 13 envs on 2 ranks (7 + 6): every rank issues the same number of gradient
    all-reduces (minibatch count from the smallest shard -- an uneven count would deadlock), the
    weights stay identical across ranks, the exploration noise differs per rank."""
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29900 + os.getpid() % 90
    procs = [ctx.Process(target=_ppo_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    gathered = out.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (w0, z0, calls0, steps0), (w1, z1, calls1, steps1) = gathered
    np.testing.assert_allclose(w0, w1, rtol=0, atol=1e-7)
    assert calls0 == calls1 and calls0 > 0
    assert z0 != z1


def test_two_rank_recurrent_ppo_with_uneven_shards_keeps_ranks_in_step():
    """The same for RecurrentPPO (LSTM policy, sequence-function learner): identical weights on both
    ranks after two iterations, same number of optimiser steps, per-rank noise."""
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29800 + os.getpid() % 90
    procs = [ctx.Process(target=_ppo_worker, args=(r, 2, port, out, True)) for r in range(2)]
    for p in procs:
        p.start()
    gathered = out.get(timeout=400)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (w0, z0, calls0, steps0), (w1, z1, calls1, steps1) = gathered
    assert len(w0) > 500_000
    np.testing.assert_allclose(w0, w1, rtol=0, atol=1e-7)
    assert calls0 == calls1 and calls0 > 0
    assert z0 != z1
