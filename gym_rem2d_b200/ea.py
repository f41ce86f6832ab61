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
sizes, so it can be spread over a process pool; evaluation itself is one ``rem2d_evaluate`` call.
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


class run2D:
    def __init__(self, config, dir, env=None, workers=0):
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
        self.generation_log = []

    # -- batched replacement of toolbox.map(toolbox.evaluate, individuals)
    def evaluate_batch(self, individuals):
        if self.env is None:
            from .env import BatchedModular2D
            self.env = BatchedModular2D()
        t0 = time.perf_counter()
        if self.workers > 1 and len(individuals) >= 4 * self.workers:
            step = (len(individuals) + self.workers - 1) // self.workers
            chunks = [(individuals[i:i + step], self.TREE_DEPTH) for i in range(0, len(individuals), step)]
            with mp.get_context("fork").Pool(self.workers) as pool:
                table = concat(pool.map(_expand_chunk, chunks))
        else:
            table = flatten_population(individuals, self.TREE_DEPTH)
        t1 = time.perf_counter()
        self.env.seed(K.TERRAIN_SEED)
        fit = self.env.evaluate(table=table, steps=self.EVALUATION_STEPS)
        t2 = time.perf_counter()
        self.last_timing = {"expand_s": t1 - t0, "evaluate_s": t2 - t1, "creature_steps": int(self.env.last_ticks.sum())}
        return [float(f) for f in fit]

    def run(self, config=None, continue_progression=False, n_generations=None):
        population = None
        if continue_progression:
            self.fitnessData = pickle.load(open(self.SAVE_FILE_DIRECTORY, "rb"))
            last = max(int(f[len("s_pop"):]) for f in os.listdir(os.path.dirname(self.SAVE_FILE_DIRECTORY)) if f.startswith("s_pop"))
            population = pickle.load(open(self.SAVE_FILE_DIRECTORY + self.POPULATION_FILE + str(last), "rb"))
        return self.run_deap(config or self.config, population=population, n_generations=n_generations)

    def run_deap(self, config, population=None, useTQDM=False, n_generations=None):
        N_GENERATIONS = 1 + int(int(config['ea']['n_evaluations']) / self.POPULATION_SIZE)
        N_GENERATIONS -= len(self.fitnessData.avg)
        if n_generations is not None:
            N_GENERATIONS = n_generations
        if population is None:
            population = [Individual.random(self.moduleList, self.config) for _ in range(self.POPULATION_SIZE)]
            for ind, fit in zip(population, self.evaluate_batch(population)):
                ind.fitness = fit
        gen = 0
        for i in range(N_GENERATIONS):
            gen += 1
            t0 = time.perf_counter()
            offspring = selTournament(population, len(population), tournsize=4)
            offspring = [copy.deepcopy(o) for o in offspring]
            for o in offspring:
                Individual.mutate(self.MORPH_MUTATION_RATE, self.MUTATION_RATE, self.MUT_SIGMA, o)
                o.fitness = 0
            fitness_values = self.evaluate_batch(offspring)
            for ind, fit in zip(offspring, fitness_values):
                ind.fitness = fit
            population = offspring                                   # no elitism, like the reference
            self.EVALUATION_NR += len(population)
            self.fitnessData.addFitnessData(fitness_values, gen)
            self.generation_log.append({"generation": i + 1, "min": float(np.min(fitness_values)), "max": float(np.max(fitness_values)),
                                        "mean": float(np.mean(fitness_values)), "seconds": time.perf_counter() - t0,
                                        **self.last_timing})
            if self.SAVEDATA:
                if i % self.CHECKPOINT_FREQUENCY == 0 or i == N_GENERATIONS:
                    self.fitnessData.save(self.SAVE_FILE_DIRECTORY)
                    pickle.dump(population, open(self.SAVE_FILE_DIRECTORY + self.POPULATION_FILE + str(i), "wb"))
                bestfit, best = 0.0, None
                for o in offspring:
                    if o.fitness > bestfit:
                        bestfit, best = o.fitness, o
                if best is not None:
                    pickle.dump(best, open(self.SAVE_FILE_DIRECTORY + self.BEST_INDIVIDUAL_FILE + str(i), "wb"))
            if time.time() - self.start_time > int(config.get("ea", "wallclock_time_limit", fallback=str(2 ** 62))):
                break
        self.population = population
        return population
