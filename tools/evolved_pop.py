"""Evaluation time of an EVOLVED-like population (config 5 after the first selection): tournament winners (size 4) of the bench
population, i.e. what one generation of run_deap evaluates when mutation is rare - many copies of long-lived creatures.
usage: python tools/evolved_pop.py ["opt=value;opt=value" ...]   (a leading PRIO gives the parents' lifetimes as priority hint)"""
import os, sys
import numpy as np
sys.path.insert(0, ".")
from gym_rem2d_b200 import constants as K, terrain
from gym_rem2d_b200.capi import Engine
from gym_rem2d_b200.population import random_population

pop = random_population(65536, ("lsystem",), seed=2, cache_dir="/tmp/rem2d_cache")
xs, ys = terrain.generate_terrain()
e = Engine(device=0); e.set_terrain(ys, K.TERRAIN_STEP)
fit, ticks = e.evaluate(pop, K.EVALUATION_STEPS)
e.close()
rng = np.random.RandomState(0)
for gen in range(2):
    asp = rng.randint(0, len(fit), size=(len(fit), 4))
    win = asp[np.arange(len(fit)), np.argmax(fit[asp], axis=1)]
    sub = pop.select(win)
    parent_ticks = ticks[win].astype(np.float32)
    print("generation %d: %d creatures, parents' mean lifetime %.1f, >=200: %d, >256: %d, max %d" % (
        gen + 1, sub.n_creatures, parent_ticks.mean(), (parent_ticks >= 200).sum(), (parent_ticks > 256).sum(), parent_ticks.max()))
    for cfg in sys.argv[1:] or [""]:
        opts = [kv for kv in cfg.split(";") if kv]
        g = Engine(device=0); g.set_terrain(ys, K.TERRAIN_STEP)
        for kv in opts:
            if kv != "PRIO":
                k_, v_ = kv.split("="); g.set_option(k_, float(v_))
        ms = []
        for _ in range(2):
            if "PRIO" in opts:
                g.set_priority(parent_ticks)
            f2, t2 = g.evaluate(sub, K.EVALUATION_STEPS)
            ms.append(g.last_step_ms())
        print("   %-48s %s ms  %.3g creature-steps/s" % (cfg or "(defaults)", " ".join("%.0f" % m for m in ms), t2.sum() / min(ms) * 1e3), flush=True)
        g.close()
    pop, fit, ticks = sub, f2, t2
