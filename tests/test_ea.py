"""Batched EA driver (8f N1): the generational loop of the reference around one batched evaluation per generation.
The loop logic is tested on CPU with a stub evaluator; the gpu-marked test runs real generations on the device."""
import os
import pickle
import random

import numpy as np
import pytest

from gym_rem2d_b200 import ea
from gym_rem2d_b200.individual import Individual


class StubEnv:
    """Deterministic stand-in for BatchedModular2D: fitness = number of bodies (no physics)."""

    def seed(self, s):
        pass

    def evaluate(self, table=None, steps=None, **kw):
        self.last_ticks = np.diff(table.body_off) * 10
        return np.diff(table.body_off).astype(np.float64)


def test_tournament_prefers_fitter_individuals():
    random.seed(0)
    class I:  # noqa: E742
        def __init__(self, f): self.fitness = f
    pop = [I(f) for f in range(100)]
    chosen = ea.selTournament(pop, 2000, 4)
    assert len(chosen) == 2000
    assert np.mean([c.fitness for c in chosen]) > 75          # E[max of 4 uniform draws] = 80


def test_generations_checkpoints_and_formats(tmp_path):
    random.seed(2)
    cfg = ea.default_config(directory=str(tmp_path), enc="lsystem", mr=0.2, mmr=0.2, ms=0.2)
    cfg["ea"]["batch_size"] = "24"
    cfg["experiment"]["checkpoint_frequency"] = "2"
    run = ea.run2D(cfg, str(tmp_path), env=StubEnv())
    pop = run.run_deap(cfg, n_generations=4)
    assert len(pop) == 24 and all(isinstance(p, Individual) for p in pop)
    assert len(run.fitnessData.avg) == 4 and run.fitnessData.p_100[-1] >= run.fitnessData.p_0[-1]
    files = sorted(os.listdir(tmp_path))
    assert "s_" in files and "s_pop0" in files and "s_pop2" in files and any(f.startswith("s_elite") for f in files)
    from gym_rem2d_b200 import refpickle
    assert "s_pop3" in files                                    # the last generation is always checkpointed
    saved = refpickle.load(tmp_path / "s_pop2")
    assert len(saved) == 24 and saved[0].genome.create(saved[0].tree_depth).getNodes()
    assert b"REM2D_main" in open(tmp_path / "s_pop2", "rb").read()      # written under the reference's class paths
    fd = refpickle.load(tmp_path / "s_")
    assert isinstance(fd, ea.FitnessData) and len(fd.avg) == 4
    # resume: continues from the newest population, absolute generation numbers, history cut to the checkpoint
    run2 = ea.run2D(cfg, str(tmp_path), env=StubEnv())
    pop2 = run2.run(cfg, continue_progression=True, n_generations=3)
    assert len(pop2) == 24 and len(run2.fitnessData.avg) == 7
    files = sorted(os.listdir(tmp_path))
    assert "s_pop4" in files and "s_pop6" in files and "s_pop5" not in files
    # resume a second time: picks s_pop6 (the newest), never an older file; nothing is overwritten out of order
    run3 = ea.run2D(cfg, str(tmp_path), env=StubEnv())
    run3.run(cfg, continue_progression=True, n_generations=2)
    assert run3.generation_offset == 7 and len(run3.fitnessData.avg) == 9
    assert [g["generation"] for g in run3.generation_log] == [8, 9]
    assert "s_pop8" in os.listdir(tmp_path)
    # selection pressure with the body-count fitness: creatures grow
    assert run.generation_log[-1]["mean"] >= run.generation_log[0]["mean"] - 1.0


def test_parallel_expansion_equals_serial():
    random.seed(3)
    cfg = ea.default_config(enc="direct")
    inds = [Individual.random(config=cfg) for _ in range(40)]
    a = ea.run2D(cfg, "", env=StubEnv(), workers=0)
    b = ea.run2D(cfg, "", env=StubEnv(), workers=4)
    assert a.evaluate_batch(inds) == b.evaluate_batch(inds)
    # variation in the persistent workers: same population size, valid offspring, tables consistent with the offspring
    random.seed(5)
    off, table = b.vary_and_expand(inds)
    assert len(off) == 40 and table.n_creatures == 40
    from gym_rem2d_b200.flatten import flatten_population
    assert np.array_equal(flatten_population(off, b.TREE_DEPTH).x0, table.x0)
    assert b.pool is not None
    b.close()


def test_packed_population_generations(tmp_path):
    """Large populations with a worker pool live in the driving process as pickles (PackedIndividual): same loop, same
    files, Individuals come back at the end; a parent that wins several tournaments yields INDEPENDENT offspring."""
    random.seed(6)
    cfg = ea.default_config(directory=str(tmp_path), enc="lsystem", mr=0.3, mmr=0.3, ms=0.3)
    cfg["ea"]["batch_size"] = "160"
    cfg["experiment"]["checkpoint_frequency"] = "2"
    run = ea.run2D(cfg, str(tmp_path), env=StubEnv(), workers=2)
    run.pipeline_halves = False                        # one evaluation per generation (the pipelined form has its own test)
    try:
        pop = run.run_deap(cfg, n_generations=3)
        assert len(pop) == 160 and all(isinstance(p, Individual) for p in pop)
        assert len({id(p) for p in pop}) == 160
        from gym_rem2d_b200.flatten import flatten_population
        nb = np.diff(flatten_population(pop, run.TREE_DEPTH).body_off)
        assert [p.fitness for p in pop] == [float(n) for n in nb]            # fitness travelled with the right individual
        assert [p.lifetime for p in pop] == [int(n) * 10 for n in nb]
        from gym_rem2d_b200 import refpickle
        saved = refpickle.load(tmp_path / "s_pop2")
        assert len(saved) == 160 and isinstance(saved[0], Individual)
        assert any(f.startswith("s_elite") for f in os.listdir(tmp_path))
        # one parent, many tournaments won: every offspring is its own clone with its own mutations
        one = ea.PackedIndividual.pack(pop[0])
        random.seed(9)
        off, table = run.vary_and_expand([one] * 160)
        assert table.n_creatures == 160 and len({o.blob for o in off}) > 100
        # a resumed run packs the loaded population again
        run2 = ea.run2D(cfg, str(tmp_path), env=StubEnv(), workers=2)
        run2.materialize_result = False
        pop2 = run2.run(cfg, continue_progression=True, n_generations=1)
        assert len(pop2) == 160 and isinstance(pop2[0], ea.PackedIndividual) and isinstance(pop2[0].unpack(), Individual)
        run2.close()
    finally:
        run.close()


def test_packed_population_two_halves_pipeline(monkeypatch):
    """Default for packed populations (REM2D_EA_PIPELINE=0 turns it off): the first half of a generation is evaluated while the
    workers expand the second half; fitness and lifetimes still land on the right individuals."""
    monkeypatch.delenv("REM2D_EA_PIPELINE", raising=False)
    random.seed(6)
    cfg = ea.default_config(enc="lsystem", mr=0.3, mmr=0.3, ms=0.3)
    cfg["ea"]["batch_size"] = "160"
    run = ea.run2D(cfg, "", env=StubEnv(), workers=2)
    run.pipeline_min = 0                               # (by default only populations of >= 65536 are split)
    try:
        assert run.pipeline_halves
        pop = run.run_deap(cfg, n_generations=3)
        from gym_rem2d_b200.flatten import flatten_population
        nb = np.diff(flatten_population(pop, run.TREE_DEPTH).body_off)
        assert len(pop) == 160 and [p.fitness for p in pop] == [float(n) for n in nb]
        assert [p.lifetime for p in pop] == [int(n) * 10 for n in nb]
        assert all(g["creature_steps"] == int(nb_.sum()) * 10 for g, nb_ in [(run.generation_log[-1], nb)])
    finally:
        run.close()


def test_repeated_parents_are_cloned_in_the_worker():
    random.seed(8)
    cfg = ea.default_config(enc="lsystem")
    a = Individual.random(config=cfg)
    off, table = ea._vary_chunk(pickle.loads(pickle.dumps(([a, a, a], 8, 0.5, 0.5, 0.5, 123))))
    assert len({id(o) for o in off}) == 3 and table.n_creatures == 3


@pytest.mark.gpu
def test_generations_on_the_gpu():
    random.seed(4)
    cfg = ea.default_config(enc="lsystem", mr=0.1, mmr=0.1, ms=0.2)
    cfg["ea"]["batch_size"] = "256"
    run = ea.run2D(cfg, "")
    pop = run.run_deap(cfg, n_generations=3)
    assert len(pop) == 256 and len(run.generation_log) == 3
    assert all(g["creature_steps"] > 256 * 40 for g in run.generation_log)
    assert run.generation_log[-1]["max"] > 5.0
