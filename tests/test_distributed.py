"""Multi-rank logic on CPU: world_size 2 over gloo. The shards are evaluated with the oracle here
(no GPU in this container); on the GPU box the same code path runs with the CUDA engine."""
import os
import random
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from gym_rem2d_b200 import Individual, constants as K, terrain
from gym_rem2d_b200.distributed import shard_indices
from gym_rem2d_b200.flatten import flatten_population


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, table, ys, ret):
    import torch.distributed as dist
    from gym_rem2d_b200.distributed import evaluate_sharded
    from oracle.oracle import OracleEngine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def make():
        e = OracleEngine()
        e.set_terrain(ys, K.TERRAIN_STEP)
        return e
    fit, steps = evaluate_sharded(table, make, 400)
    ret[rank] = (fit, steps)
    dist.destroy_process_group()


def test_shards_partition_population_and_balance_sizes():
    random.seed(3)
    table = flatten_population([Individual.random(encoding="lsystem") for _ in range(101)])
    nb = np.diff(table.body_off)
    for world in (1, 2, 4, 8):
        parts = [shard_indices(table.body_off, r, world) for r in range(world)]
        allidx = np.sort(np.concatenate(parts))
        assert np.array_equal(allidx, np.arange(101))
        loads = [nb[p].sum() for p in parts]
        assert max(loads) - min(loads) <= nb.max() + 1


def test_two_ranks_gloo_gather_equals_single_rank():
    random.seed(4)
    table = flatten_population([Individual.random(encoding="direct") for _ in range(37)])
    xs, ys = terrain.generate_terrain()
    from oracle.oracle import OracleEngine
    e = OracleEngine()
    e.set_terrain(ys, K.TERRAIN_STEP)
    ref, ticks = e.evaluate(table, 400)
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, table, ys, ret), nprocs=2, join=True)
    for r in range(2):
        fit, steps = ret[r]
        assert np.array_equal(fit, ref.astype(np.float32))      # bitwise identical, any shard count
    assert ret[0][1] + ret[1][1] == int(ticks.sum())


def _serve_worker(rank, world, port, table, ys, ret):
    """rank 0 drives two 'generations' (table broadcast + sharded evaluation + fitness gather) and sends the stop signal; the other
    rank sits in serve_evaluations - the multi-GPU form of the EA loop (ea.run2D(distributed=True) / bench.py --ea)."""
    import torch.distributed as dist
    from gym_rem2d_b200 import distributed as rdist
    from oracle.oracle import OracleEngine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    e = OracleEngine()
    e.set_terrain(ys, K.TERRAIN_STEP)
    if rank == 0:
        out = []
        for gen in range(2):
            sub = table if gen == 0 else table.select(np.arange(table.n_creatures)[::-1].copy())
            hint = np.arange(sub.n_creatures, dtype=np.float32) * 20 if gen == 1 else None      # scheduling hint: results unchanged
            fit, lifetimes = rdist.evaluate_broadcast(sub, e, 300, gather_ticks=True, expected_ticks=hint)
            out.append((fit, lifetimes))
        assert rdist.broadcast_table(None, 0) is None          # stop signal
        ret[0] = out
    else:
        ret[rank] = rdist.serve_evaluations(e, 300, gather_ticks=True)
    dist.destroy_process_group()


def test_rank0_drives_and_the_other_ranks_serve():
    random.seed(6)
    table = flatten_population([Individual.random(encoding="lsystem") for _ in range(29)])
    xs, ys = terrain.generate_terrain()
    from oracle.oracle import OracleEngine
    e = OracleEngine()
    e.set_terrain(ys, K.TERRAIN_STEP)
    ref, ref_ticks = e.evaluate(table, 300)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_serve_worker, args=(2, _free_port(), table, ys, ret), nprocs=2, join=True)
    assert ret[1] == 2                                           # two generations served, then the stop signal
    assert np.array_equal(ret[0][0][0], ref.astype(np.float32))
    assert np.array_equal(ret[0][1][0], ref.astype(np.float32)[::-1])
    assert np.array_equal(ret[0][0][1], ref_ticks) and np.array_equal(ret[0][1][1], ref_ticks[::-1])     # gathered lifetimes
