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


def evaluate_sharded(table, engine, max_ticks, rank=None, world=None, device=None, gather_ticks=False, expected_ticks=None):
    """Evaluate this rank's shard of ``table`` on ``engine`` and gather everyone's fitness: the multi-GPU form of
    ``pool.map(evaluate, population, chunksize=ceil(pop/n))`` (REM2D_main.py:256-262). ``engine`` is a long-lived
    ``capi.Engine`` (its device buffers are grow-only and reused from generation to generation); a zero-argument factory is
    accepted for one-off calls. Returns (fitness of the WHOLE population in population order, float32; this rank's
    creature-steps). With ``gather_ticks`` the lifetimes are gathered as well (one more all_gather) and the second value is the
    lifetime of every creature of the whole population (int32) - every rank must pass the same flag."""
    import torch.distributed as dist
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
        world = dist.get_world_size() if dist.is_initialized() else 1
    sub, idx = shard_population(table, rank, world)
    eng = engine() if callable(engine) else engine
    if expected_ticks is not None:          # scheduling hint (expected lifetimes of the WHOLE population): this shard's part
        eng.set_priority(np.asarray(expected_ticks, np.float32)[idx])
    fit, ticks = eng.evaluate(sub, max_ticks)
    if gather_ticks:                       # lifetimes <= max_ticks are exact in float32
        return (gather_fitness(fit, idx, table.n_creatures, device=device),
                gather_fitness(ticks, idx, table.n_creatures, device=device).astype(np.int32))
    return gather_fitness(fit, idx, table.n_creatures, device=device), int(ticks.sum())


# ---------------------------------------------------------------------------------------------------------------------
# Rank 0 drives, the other ranks serve: the evolutionary loop (selection, variation, expansion) runs on rank 0 only; per
# generation it broadcasts the flattened table, every rank evaluates its shard, the fitness vector is all-gathered.
_FIELDS = ("body_off", "shape", "hx", "hy", "x0", "y0", "a0", "node_index", "type_ref", "joint_parent", "anchor_a", "anchor_b",
           "lower", "upper", "max_torque", "ctrl")


def broadcast_table(table, src=0, device=None):
    """Broadcast a PopulationTable from rank ``src`` (``table`` may be None elsewhere). One small header broadcast with the
    array sizes, then one broadcast per array (the whole table of 65536 creatures is ~45 MB: milliseconds over NVLink)."""
    import torch
    import torch.distributed as dist
    from .flatten import PopulationTable
    rank = dist.get_rank()
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    ref = {k: None for k in _FIELDS}
    if rank == src and table is not None:
        ref = {k: np.ascontiguousarray(getattr(table, k)) for k in _FIELDS}
    head = torch.zeros(len(_FIELDS) + 1, dtype=torch.int64, device=dev)
    if rank == src and table is not None:
        head[:-1] = torch.tensor([ref[k].size for k in _FIELDS], dtype=torch.int64)
        head[-1] = 1
    dist.broadcast(head, src)
    if int(head[-1]) == 0:
        return None                                   # stop signal
    proto = PopulationTable(*(np.zeros((0, 2) if k in ("anchor_a", "anchor_b") else ((0, 5) if k == "ctrl" else 0),
                                           dt) for k, dt in zip(_FIELDS, _DTYPES)))
    out = []
    for k, dt, n in zip(_FIELDS, _DTYPES, head[:-1].tolist()):
        if rank == src:
            t = torch.from_numpy(ref[k].reshape(-1).view(np.uint8)).to(dev)
        else:
            t = torch.empty(int(n) * np.dtype(dt).itemsize, dtype=torch.uint8, device=dev)
        dist.broadcast(t, src)
        a = t.cpu().numpy().view(dt)
        shape = getattr(proto, k).shape
        out.append(a.reshape((-1,) + shape[1:]) if len(shape) > 1 else a)
    return PopulationTable(*out)


_DTYPES = (np.int32, np.uint8, np.float32, np.float32, np.float32, np.float32, np.float32, np.int32, np.int16, np.int16, np.float32,
           np.float32, np.float32, np.float32, np.float32, np.float64)


def evaluate_broadcast(table, engine, max_ticks, device=None, gather_ticks=False, expected_ticks=None):
    """Collective: rank 0 passes the generation's table (None = stop) and, optionally, the expected lifetimes of its creatures
    (the scheduling hint of rem2d_set_priority; broadcast with the table, zeros = no hint); the others pass None; everybody
    returns the fitness of the whole population (or None on stop) and what evaluate_sharded returns second."""
    import torch
    import torch.distributed as dist
    table = broadcast_table(table, 0, device)
    if table is None:
        return None, 0
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    hint = torch.zeros(table.n_creatures, dtype=torch.float32, device=dev)
    if dist.get_rank() == 0 and expected_ticks is not None and len(expected_ticks) == table.n_creatures:
        hint.copy_(torch.as_tensor(np.asarray(expected_ticks, np.float32)))
    dist.broadcast(hint, 0)
    hint = hint.cpu().numpy()
    return evaluate_sharded(table, engine, max_ticks, dist.get_rank(), dist.get_world_size(), device, gather_ticks,
                            hint if hint.any() else None)


def serve_evaluations(engine, max_ticks, device=None, gather_ticks=False):
    """Body of every rank != 0 while rank 0 runs the evolutionary loop: evaluate shards until rank 0 sends the stop signal.
    ``gather_ticks`` as rank 0 passes it (ea.run2D: True)."""
    n = 0
    while True:
        fit, _ = evaluate_broadcast(None, engine, max_ticks, device, gather_ticks)
        if fit is None:
            return n
        n += 1
