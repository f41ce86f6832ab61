"""Synthetic random populations: ``random.seed(S)`` then repeated
``Individual.random(encoding=...)``, expanded and flattened. Generation is host-bound Python
(SURVEY.md 7.3), so large populations are produced by a process pool, one seeded chunk per task."""
import multiprocessing as mp
import os
import random

import numpy as np

from .flatten import PopulationTable, flatten_population
from .individual import Individual

CHUNK = 512


def _chunk(args):
    seed, chunk_index, n, encodings = args
    random.seed(seed * 1000003 + chunk_index)
    np.random.seed((seed * 1000003 + chunk_index) % (2 ** 32))
    inds = [Individual.random(encoding=encodings[i % len(encodings)]) for i in range(n)]
    return flatten_population(inds)


def concat(tables):
    nb = np.concatenate([np.diff(t.body_off) for t in tables])
    boff = np.zeros(len(nb) + 1, np.int32)
    np.cumsum(nb, out=boff[1:])
    cat = lambda k: np.concatenate([getattr(t, k) for t in tables])
    return PopulationTable(boff, *(cat(k) for k in ("shape", "hx", "hy", "x0", "y0", "a0", "node_index", "type_ref",
                                                     "joint_parent", "anchor_a", "anchor_b", "lower", "upper",
                                                     "max_torque", "ctrl")))


def random_population(n, encodings=("lsystem",), seed=0, workers=None, cache_dir=None):
    """``n`` random individuals, encodings cycled per individual. Deterministic in (n, encodings, seed)."""
    if isinstance(encodings, str):
        encodings = (encodings,)
    key = "rem2d_pop_%s_%d_%d.npz" % ("-".join(encodings), n, seed)
    if cache_dir:
        path = os.path.join(cache_dir, key)
        if os.path.exists(path):
            z = np.load(path)
            return PopulationTable(*(z[k] for k in ("body_off", "shape", "hx", "hy", "x0", "y0", "a0", "node_index",
                                                    "type_ref", "joint_parent", "anchor_a", "anchor_b", "lower", "upper",
                                                    "max_torque", "ctrl")))
    tasks = [(seed, i, min(CHUNK, n - i * CHUNK), tuple(encodings)) for i in range((n + CHUNK - 1) // CHUNK)]
    workers = workers or min(len(tasks), os.cpu_count() or 1)
    if workers > 1 and len(tasks) > 1:
        with mp.get_context("fork").Pool(workers) as pool:
            tables = pool.map(_chunk, tasks)
    else:
        tables = [_chunk(t) for t in tasks]
    pop = concat(tables)
    if cache_dir:
        os.makedirs(cache_dir, exist_ok=True)
        tmp = os.path.join(cache_dir, "%s.%d.tmp.npz" % (key, os.getpid()))
        np.savez(tmp, **{k: getattr(pop, k) for k in ("body_off", "shape", "hx", "hy", "x0", "y0", "a0", "node_index",
                                                     "type_ref", "joint_parent", "anchor_a", "anchor_b", "lower",
                                                     "upper", "max_torque", "ctrl")})
        os.replace(tmp, os.path.join(cache_dir, key))
    return pop
