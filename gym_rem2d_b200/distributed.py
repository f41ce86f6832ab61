"""Sharding of a population across ranks (one process per GPU) and the single collective of the path.

The reference parallelises with ``multiprocessing.Pool(n_cores).map(evaluate, population, chunksize=ceil(pop/n))``
(REM2D_main.py:256-262): static contiguous chunks, pickled individuals in, floats out. Creatures never
interact (one b2World each, Modular2DEnv.py:572), so here the flattened table is sharded by individual
with no data-path collective; only the fitness vector is all-gathered (4 bytes per creature) —
``torch.distributed`` NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
import numpy as np


def shard_indices(body_off, rank, world):
    """Indices of the creatures rank ``rank`` evaluates. Creatures are dealt round-robin in order of
    decreasing body count, so every rank gets the same mix of sizes (cost grows with the body count)
    instead of the reference's contiguous chunks."""
    nb = np.diff(np.asarray(body_off))
    order = np.argsort(-nb, kind="stable")
    return np.sort(order[rank::world])


def shard_population(table, rank, world):
    idx = shard_indices(table.body_off, rank, world)
    return table.select(idx), idx


def gather_fitness(local_fitness, local_idx, n_total, device=None):
    """all_gather of the per-rank fitness vectors into population order. Shards may differ in length by
    one, so they are padded to the longest before the collective."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        out = np.empty(n_total, np.float32)
        out[local_idx] = local_fitness
        return out
    per = (n_total + world - 1) // world
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    fit = torch.full((per,), float("nan"), dtype=torch.float32, device=dev)
    idx = torch.full((per,), -1, dtype=torch.int64, device=dev)
    fit[: len(local_idx)] = torch.as_tensor(np.asarray(local_fitness, np.float32), device=dev)
    idx[: len(local_idx)] = torch.as_tensor(np.asarray(local_idx, np.int64), device=dev)
    fit_all = torch.empty(world * per, dtype=torch.float32, device=dev)
    idx_all = torch.empty(world * per, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(fit_all, fit)
    dist.all_gather_into_tensor(idx_all, idx)
    fit_all, idx_all = fit_all.cpu().numpy(), idx_all.cpu().numpy()
    out = np.empty(n_total, np.float32)
    m = idx_all >= 0
    out[idx_all[m]] = fit_all[m]
    return out


def evaluate_sharded(table, engine, max_ticks, rank=None, world=None, device=None):
    """Evaluate this rank's shard of ``table`` on ``engine`` and gather everyone's fitness: the multi-GPU form of
    ``pool.map(evaluate, population, chunksize=ceil(pop/n))`` (REM2D_main.py:256-262). ``engine`` is a long-lived
    ``capi.Engine`` (its device buffers are grow-only and reused from generation to generation); a zero-argument factory is
    accepted for one-off calls. Returns (fitness of the WHOLE population in population order, float32; this rank's
    creature-steps)."""
    import torch.distributed as dist
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
        world = dist.get_world_size() if dist.is_initialized() else 1
    sub, idx = shard_population(table, rank, world)
    eng = engine() if callable(engine) else engine
    fit, ticks = eng.evaluate(sub, max_ticks)
    return gather_fitness(fit, idx, table.n_creatures, device=device), int(ticks.sum())
