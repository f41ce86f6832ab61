"""Batched evolutionary driver: the reference's generational loop with the per-individual
``toolbox.map(toolbox.evaluate, offspring)`` replaced by ONE batched GPU evaluation per generation.

Restates ``run2D.run_deap`` (REM2D_main.py:241-348) without DEAP (not a dependency here): tournament selection of
size 4 on ``ind.fitness`` with replacement (``deap.tools.selTournament`` semantics: k tournaments, each drawing
``tournsize`` aspirants with ``random.choice``), deep-copied offspring, ``Individual.mutate``, no elitism
(``population = offspring``), per-generation fitness percentiles (``DataAnalysis.FitnessData``, DataAnalysis.py:39-56),
pickled checkpoints ``<out>/s_pop<i>``, ``<out>/s_`` and ``<out>/s_elite<i>`` every ``checkpoint_frequency``
generations / on every generation with a positive best (REM2D_main.py:192-195,311-329). The INI keys are the
reference's (Experiments/configuration_maker.py:10-63, 0.cfg).

Host-side expansion (mutate -> genome.create -> flatten) is Python and dominates a generation at large population
sizes, so variation AND expansion run in a PERSISTENT process pool (created in ``__init__``, i.e. before
the CUDA context exists in this process; ``forkserver`` start method, so no worker ever inherits driver state): the parent
only draws the tournament winners and ships chunks of parents, the workers clone, mutate and flatten them and return
(offspring, table). Large populations live in the driving process as pickles (PackedIndividual), so that no individual is
ever (un)pickled there. Evaluation is one ``rem2d_evaluate`` call per generation (two halves on a single device, the
second half expanding while the first is evaluated), or - under ``torch.distributed`` - one call per rank on its shard of
the table plus all_gathers of the fitness and lifetime vectors (distributed.evaluate_broadcast).
Checkpoints are written under the reference's class paths (refpickle.py) and numbered by ABSOLUTE generation, so a resumed
run never overwrites or re-reads an older population.
"""
import configparser
import copy
import multiprocessing as mp
import os
import pickle
import random
import time

import numpy as np

from . import constants as K
from .flatten import flatten_population
from .individual import Individual, get_module_list
from .population import concat


def default_config(directory="", enc="lsystem", mr=0.01, mmr=0.01, ms=0.1):
    """The reference's auto-generated configuration (configuration_maker.create)."""
    c = configparser.ConfigParser()
    c["experiment"] = {"checkpoint_frequency": "10", "save_elite": "1", "experiment_number": "0", "directory": directory}
    c["ea"] = {"n_evaluations": "10000", "batch_size": "100", "mutation_prob": str(mr), "morphmutation_prob": str(mmr),
               "mutation_sigma": str(ms), "headless": "1", "show_best": "0", "load_best": "0", "n_cores": "6", "interval": "5",
               "wallclock_time_limit": str(2 ** 62)}
    c["morphology"] = {"max_size": "40", "max_depth": "7", "m_rectangle": "4", "m_circular": "4"}
    c["evaluation"] = {"wod_speed": "2"}
    c["encoding"] = {"type": enc}
    c["control"] = {"type": "wave"}
    c["visualization"] = {"v_tree": "0", "v_progression": "0", "v_debug": "0"}
    return c


class FitnessData:
    """Progress container pickled to ``<out>/s_`` (same attributes as DataAnalysis.FitnessData)."""

    def __init__(self):
        self.p_0, self.p_25, self.p_50, self.p_75, self.p_100, self.avg, self.divValues = [], [], [], [], [], [], []

    def save(self, saveFile, num=''):
        pickle.dump(self, open(saveFile + str(num), "wb"))

    def addFitnessData(self, fitnesses, gen):
        self.avg.append(np.average(fitnesses))
        for p in (0, 25, 50, 75, 100):
            getattr(self, "p_%d" % p).append(np.percentile(fitnesses, p))


def selTournament(individuals, k, tournsize, fit_attr="fitness"):
    chosen = []
    for _ in range(k):
        aspirants = [random.choice(individuals) for _ in range(tournsize)]
        chosen.append(max(aspirants, key=lambda ind: getattr(ind, fit_attr)))
    return chosen


def _expand_chunk(args):
    inds, depth = args
    return flatten_population(inds, depth)


def _random_chunk(args):
    """Worker: ``n`` random individuals (REM2D_main.py:296-297 creates the initial population one by one)."""
    n, module_list, config, seed = args
    random.seed(seed)
    np.random.seed(seed % (2 ** 32))
    return [Individual.random(module_list, config) for _ in range(n)]


class PackedIndividual:
    """An individual kept as its pickle. Between two generations the driving process needs of an individual only its fitness
    (tournaments) and its lifetime (scheduling hint); the object itself is needed by the WORKER that varies it. Shipping objects
    costs the driving process a serial pickle.dumps per parent and a serial pickle.loads per offspring - measured 0.17 + 0.15 ms
    per individual, 7-19 s per generation at population 65536, more than everything else together - while bytes go through
    the pipe as one copy. ``unpack()`` gives the Individual (checkpoints, the elite, the result of the run)."""
    __slots__ = ("blob", "fitness", "lifetime")

    def __init__(self, blob, fitness=0.0, lifetime=0):
        self.blob, self.fitness, self.lifetime = blob, fitness, lifetime

    @classmethod
    def pack(cls, ind):
        return cls(pickle.dumps(ind, pickle.HIGHEST_PROTOCOL), getattr(ind, "fitness", 0.0), getattr(ind, "lifetime", 0))

    def unpack(self):
        ind = pickle.loads(self.blob)
        ind.fitness, ind.lifetime = self.fitness, self.lifetime
        return ind


def _random_chunk_packed(args):
    """Worker: ``n`` random individuals, returned as pickles + their flattened table."""
    inds = _random_chunk(args[:4])
    return [pickle.dumps(o, pickle.HIGHEST_PROTOCOL) for o in inds], flatten_population(inds, args[4])


def _vary_chunk_packed(args):
    """Worker: like _vary_chunk on pickles. ``blobs`` holds every distinct parent of the chunk once, ``idx`` the chunk's
    parents as indices into it; every occurrence is unpickled separately (= ``toolbox.clone``)."""
    blobs, idx, depth, mmr, mr, sigma, seed = args
    random.seed(seed)
    np.random.seed(seed % (2 ** 32))
    parents = [pickle.loads(blobs[i]) for i in idx]
    for o in parents:
        Individual.mutate(mmr, mr, sigma, o)
        o.fitness = 0
    table = flatten_population(parents, depth)
    return [pickle.dumps(o, pickle.HIGHEST_PROTOCOL) for o in parents], table


def _vary_chunk(args):
    """Worker: variation + expansion of one chunk of selected parents (REM2D_main.py:283-290 + evaluate's genome.create).
    Parents arrive pickled (= the deep copy of ``toolbox.clone``); every chunk is seeded so that a run is reproducible
    for a fixed number of workers."""
    parents, depth, mmr, mr, sigma, seed = args
    random.seed(seed)
    np.random.seed(seed % (2 ** 32))
    seen = set()
    for k, o in enumerate(parents):          # a parent that won several tournaments arrives as ONE object: clone the repeats
        if id(o) in seen:
            parents[k] = copy.deepcopy(o)
        seen.add(id(o))
    for o in parents:
        Individual.mutate(mmr, mr, sigma, o)
        o.fitness = 0
    return parents, flatten_population(parents, depth)


class run2D:
    def __init__(self, config, dir, env=None, workers=0, distributed=False):
        self.config = config
        self.start_time = time.time()
        self.fitnessData = FitnessData()
        self.BEST_INDIVIDUAL_FILE, self.POPULATION_FILE = "elite", "pop"
        self.SAVE_FILE_DIRECTORY = os.path.join(dir, 's_')
        self.CHECKPOINT_FREQUENCY = int(config['experiment']['checkpoint_frequency'])
        self.POPULATION_SIZE = int(config['ea']['batch_size'])
        self.MUTATION_RATE = float(config['ea']['mutation_prob'])
        self.MORPH_MUTATION_RATE = float(config['ea']['morphmutation_prob'])
        self.MUT_SIGMA = float(config['ea']['mutation_sigma'])
        self.TREE_DEPTH = int(config['morphology']['max_depth'])
        self.EVALUATION_STEPS = K.EVALUATION_STEPS
        self.SAVEDATA = bool(dir)
        self.EVALUATION_NR = 0
        self.moduleList = get_module_list()
        self.env = env
        self.workers = workers
        self.distributed = distributed       # evaluate this rank's shard and all_gather (torch.distributed must be initialised)
        self.generation_log = []
        self.generation_offset = 0           # generations already done by the run this one resumes
        self.expected_ticks = None           # scheduling hint for the next evaluation (parents' lifetimes)
        self.last_lifetimes = None
        self.last_device_s = 0.0
        # packed populations: evaluate the first half of a generation while the workers expand the second half. Measured at
        # population 65536 (1 GPU, 14 workers): 4.33 s per generation against 4.93 s - two evaluations of 32768 cost 0.3-0.5 s
        # more than one of 65536, the overlap hides 0.8-1.1 s of expansion (REM2D_EA_PIPELINE=0 turns it off)
        self.pipeline_halves = os.environ.get("REM2D_EA_PIPELINE", "1") != "0"
        self.pipeline_min = 65536            # smallest population that is evaluated in two halves (measured at 65536 only)
        self.materialize_result = True       # run_deap returns Individuals (False: PackedIndividuals where the population was packed)
        # persistent workers, started BEFORE any CUDA work of this process (the engine is created lazily, later)
        # (forkserver re-imports __main__ in the workers, which an interactive / stdin main cannot offer: plain fork there - still
        # before any CUDA work of this process)
        import sys
        method = "forkserver" if getattr(sys.modules.get("__main__"), "__file__", None) else "fork"
        self.pool = mp.get_context(method).Pool(workers) if workers > 1 else None

    def close(self):
        if self.pool is not None:
            self.pool.terminate()
            self.pool = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chunks(self, items):
        step = max(1, (len(items) + max(self.workers, 1) * 4 - 1) // (max(self.workers, 1) * 4))
        return [items[i:i + step] for i in range(0, len(items), step)]

    def _ensure_env(self):
        if self.env is None:
            from .env import BatchedModular2D
            device = 0
            if self.distributed:
                device = int(os.environ.get("LOCAL_RANK", "0"))
            self.env = BatchedModular2D(device=device)
        return self.env

    # -- batched replacement of toolbox.map(toolbox.evaluate, individuals)
    def evaluate_table(self, table):
        env = self._ensure_env()
        env.seed(K.TERRAIN_SEED)
        if self.distributed:
            # rank 0 drives the loop and broadcasts the generation's table; the other ranks sit in distributed.serve_evaluations
            from . import distributed as rdist
            hint = self.expected_ticks if self.expected_ticks is not None and len(self.expected_ticks) == table.n_creatures else None
            fit, lifetimes = rdist.evaluate_broadcast(table, env.engine, self.EVALUATION_STEPS, gather_ticks=True, expected_ticks=hint)
            self.last_lifetimes = lifetimes
            return [float(f) for f in fit], int(lifetimes.sum())
        if self.expected_ticks is not None and len(self.expected_ticks) == table.n_creatures and hasattr(env, "engine"):
            env.engine.set_priority(self.expected_ticks)        # offspring are expected to live about as long as their parents
        fit = env.evaluate(table=table, steps=self.EVALUATION_STEPS)
        self.last_lifetimes = np.asarray(env.last_ticks)
        if hasattr(env, "engine"):
            self.last_device_s = env.engine.last_step_ms() * 1e-3      # the episode launches alone (CUDA events)
        return [float(f) for f in fit], int(env.last_ticks.sum())

    def evaluate_batch(self, individuals):
        t0 = time.perf_counter()
        if self.pool is not None and len(individuals) >= 4 * self.workers:
            table = concat(self.pool.map(_expand_chunk, [(c, self.TREE_DEPTH) for c in self._chunks(individuals)]))
        else:
            table = flatten_population(individuals, self.TREE_DEPTH)
        t1 = time.perf_counter()
        fit, steps = self.evaluate_table(table)
        t2 = time.perf_counter()
        self.last_timing = {"expand_s": t1 - t0, "evaluate_s": t2 - t1, "creature_steps": steps}
        return fit

    def vary_expand_evaluate(self, parents):
        """One generation's variation, expansion and evaluation. Packed populations on a single device are PIPELINED in two
        halves: the workers return (offspring pickles, table) chunk by chunk in order, the first half of the chunks is evaluated
        on the GPU while the workers expand the second half. (A finer-grained form - evaluate every few thousand creatures as
        they arrive - was measured and dropped: many small evaluations each pay the lifetime of their longest-lived creature,
        6-8 s per generation instead of 1 s for one evaluation, profiles/r2_ea_config5.json.) Returns (offspring, fitness
        list, timing)."""
        t0 = time.perf_counter()
        packed = bool(parents) and isinstance(parents[0], PackedIndividual)
        # (single device only: on 2 GPUs the two half-size collective evaluations cost what the overlap hides - 3.45 vs 3.47 s)
        # and only for populations whose halves still fill the GPU: a latency-bound evaluation takes one creature lifetime
        # whatever its size (8192 creatures: 0.57 s for two halves against 0.24 s for one evaluation)
        if packed and self.pool is not None and self.pipeline_halves and not self.distributed and len(parents) >= self.pipeline_min:
            # two halves: the first half of the chunks is evaluated while the workers expand the second half
            jobs = self._packed_jobs(parents)
            it = self.pool.imap(_vary_chunk_packed, jobs)
            half = max(1, len(jobs) // 2)
            expected, offspring, fit, lifetimes, tables, steps, eval_s, dev_s = self.expected_ticks, [], [], [], [], 0, 0.0, 0.0
            for lo, hi in ((0, half), (half, len(jobs))):
                parts = [next(it) for _ in range(lo, hi)]
                if not parts:
                    continue
                table = concat([t for _, t in parts])
                n0 = len(offspring)
                offspring.extend(PackedIndividual(b) for bl, _ in parts for b in bl)
                if expected is not None and len(expected) == len(parents):
                    self.expected_ticks = expected[n0:n0 + table.n_creatures]
                te = time.perf_counter()
                f, s_ = self.evaluate_table(table)
                eval_s += time.perf_counter() - te
                dev_s += self.last_device_s
                fit.extend(f); steps += s_; lifetimes.append(self.last_lifetimes); tables.append(table)
            self.expected_ticks = expected
            self.last_lifetimes = np.concatenate(lifetimes)
            total = time.perf_counter() - t0
            return offspring, fit, {"expand_s": total - eval_s, "evaluate_s": eval_s, "creature_steps": steps,
                                    "evaluate_device_s": dev_s,
                                    "mean_bodies": float(np.concatenate([np.diff(t.body_off) for t in tables]).mean())}
        offspring, table = self.vary_and_expand(parents)
        t1 = time.perf_counter()
        fit, steps = self.evaluate_table(table)
        t2 = time.perf_counter()
        return offspring, fit, {"expand_s": t1 - t0, "evaluate_s": t2 - t1, "creature_steps": steps,
                                "evaluate_device_s": self.last_device_s, "mean_bodies": float(np.diff(table.body_off).mean())}

    def _packed_jobs(self, parents):
        """Worker jobs for packed parents: every distinct parent of a chunk once as bytes + the chunk as indices into them."""
        jobs = []
        for chunk in self._chunks(parents):
            slot, blobs, idx = {}, [], []
            for p in chunk:
                k = slot.get(id(p))
                if k is None:
                    k = slot[id(p)] = len(blobs)
                    blobs.append(p.blob)
                idx.append(k)
            jobs.append((blobs, idx, self.TREE_DEPTH, self.MORPH_MUTATION_RATE, self.MUTATION_RATE, self.MUT_SIGMA,
                         random.getrandbits(48)))
        return jobs

    def vary_and_expand(self, parents):
        """clone + mutate + expand: in the workers when there is a pool, else here."""
        if self.pool is not None and parents and isinstance(parents[0], PackedIndividual):
            parts = self.pool.map(_vary_chunk_packed, self._packed_jobs(parents))
            return [PackedIndividual(b) for bl, _ in parts for b in bl], concat([t for _, t in parts])
        if self.pool is not None and len(parents) >= 4 * self.workers:
            jobs = [(c, self.TREE_DEPTH, self.MORPH_MUTATION_RATE, self.MUTATION_RATE, self.MUT_SIGMA, random.getrandbits(48))
                    for c in self._chunks(parents)]
            parts = self.pool.map(_vary_chunk, jobs)
            return [o for p, _ in parts for o in p], concat([t for _, t in parts])
        offspring = [copy.deepcopy(o) for o in parents]
        for o in offspring:
            Individual.mutate(self.MORPH_MUTATION_RATE, self.MUTATION_RATE, self.MUT_SIGMA, o)
            o.fitness = 0
        return offspring, flatten_population(offspring, self.TREE_DEPTH)

    def run(self, config=None, continue_progression=False, n_generations=None):
        from . import refpickle
        population = None
        if continue_progression:
            self.fitnessData = refpickle.load(self.SAVE_FILE_DIRECTORY)
            d = os.path.dirname(self.SAVE_FILE_DIRECTORY)
            last = max(int(f[len("s_pop"):]) for f in os.listdir(d) if f.startswith("s_pop") and f[len("s_pop"):].isdigit())
            population = refpickle.load(self.SAVE_FILE_DIRECTORY + self.POPULATION_FILE + str(last))
            # s_ may hold generations after the newest population checkpoint: the run continues from the population,
            # so the fitness history is cut back to it (absolute generation = index of the checkpoint + 1)
            self.generation_offset = last + 1
            for k in ("p_0", "p_25", "p_50", "p_75", "p_100", "avg"):
                setattr(self.fitnessData, k, list(getattr(self.fitnessData, k))[:self.generation_offset])
        return self.run_deap(config or self.config, population=population, n_generations=n_generations)

    @staticmethod
    def _unpacked(population):
        return [p.unpack() if isinstance(p, PackedIndividual) else p for p in population]

    def _checkpoint(self, population, g):
        from . import refpickle
        refpickle.dump(self.fitnessData, self.SAVE_FILE_DIRECTORY)
        refpickle.dump(self._unpacked(population), self.SAVE_FILE_DIRECTORY + self.POPULATION_FILE + str(g))

    def run_deap(self, config, population=None, useTQDM=False, n_generations=None):
        from . import refpickle
        N_GENERATIONS = 1 + int(int(config['ea']['n_evaluations']) / self.POPULATION_SIZE)
        N_GENERATIONS -= len(self.fitnessData.avg)
        if n_generations is not None:
            N_GENERATIONS = n_generations
        rank0 = (not self.distributed) or int(os.environ.get("RANK", "0")) == 0
        # large populations with a worker pool: the population lives here as pickles (see PackedIndividual)
        packed = self.pool is not None and self.POPULATION_SIZE >= 64 * self.workers
        if population is None:
            if packed:
                sizes = [len(c) for c in self._chunks(list(range(self.POPULATION_SIZE)))]
                parts = self.pool.map(_random_chunk_packed, [(n, self.moduleList, self.config, random.getrandbits(48), self.TREE_DEPTH)
                                                             for n in sizes])
                population = [PackedIndividual(b) for bl, _ in parts for b in bl]
                fitness_values, _ = self.evaluate_table(concat([t for _, t in parts]))
            else:
                population = [Individual.random(self.moduleList, self.config) for _ in range(self.POPULATION_SIZE)]
                fitness_values = self.evaluate_batch(population)
            for k, (ind, fit) in enumerate(zip(population, fitness_values)):
                ind.fitness = fit
                if self.last_lifetimes is not None and len(self.last_lifetimes) == len(population):
                    ind.lifetime = int(self.last_lifetimes[k])
        elif packed:
            population = [p if isinstance(p, PackedIndividual) else PackedIndividual.pack(p) for p in population]
        for i in range(N_GENERATIONS):
            g = self.generation_offset + i                     # absolute generation index (file suffix)
            t0 = time.perf_counter()
            parents = selTournament(population, len(population), tournsize=4)
            self.expected_ticks = np.array([getattr(p, "lifetime", 0) for p in parents], np.float32)
            offspring, fitness_values, timing = self.vary_expand_evaluate(parents)
            for k, (ind, fit) in enumerate(zip(offspring, fitness_values)):
                ind.fitness = fit
                if self.last_lifetimes is not None and len(self.last_lifetimes) == len(offspring):
                    ind.lifetime = int(self.last_lifetimes[k])
            population = offspring                                   # no elitism, like the reference
            self.EVALUATION_NR += len(population)
            self.fitnessData.addFitnessData(fitness_values, g + 1)
            self.generation_log.append({"generation": g + 1, "min": float(np.min(fitness_values)), "max": float(np.max(fitness_values)),
                                        "mean": float(np.mean(fitness_values)), "seconds": time.perf_counter() - t0, **timing})
            if self.SAVEDATA and rank0:
                if g % self.CHECKPOINT_FREQUENCY == 0 or i == N_GENERATIONS - 1:     # the last generation is always saved
                    self._checkpoint(population, g)
                bestfit, best = 0.0, None
                for o in offspring:
                    if o.fitness > bestfit:
                        bestfit, best = o.fitness, o
                if best is not None:
                    refpickle.dump(self._unpacked([best])[0], self.SAVE_FILE_DIRECTORY + self.BEST_INDIVIDUAL_FILE + str(g))
            if time.time() - self.start_time > int(config.get("ea", "wallclock_time_limit", fallback=str(2 ** 62))):
                break
        if self.materialize_result:
            population = self._unpacked(population)
        self.population = population
        return population
