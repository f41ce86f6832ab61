"""GPU experiment: whole-population evaluation time of small / mid-size populations for different execution strategies:
queue mode with one lane per creature (G=1), queue mode with automatic group widths, forced group widths, and a warp per
creature from tick 0."""
import os, sys
import numpy as np
sys.path.insert(0, ".")
from gym_rem2d_b200 import constants as K, terrain
from gym_rem2d_b200.capi import Engine
from gym_rem2d_b200.population import random_population

big = random_population(32768, ("lsystem",), seed=2, cache_dir="/tmp/rem2d_cache")
xs, ys = terrain.generate_terrain()
MODES = (("G=1", {"warp_mode_max": 0, "group_shift": 0}), ("auto", {"warp_mode_max": 0}), ("G=4", {"warp_mode_max": 0, "group_shift": 2}),
         ("G=8", {"warp_mode_max": 0, "group_shift": 3}), ("G=16", {"warp_mode_max": 0, "group_shift": 4}), ("warp", {"warp_mode_max": 1e9}))
for n in [int(a) for a in sys.argv[1:]] or (128, 1024, 4096, 8192, 16384, 32768):
    pop = big.select(np.arange(n))
    res = {}
    for mode, opts in MODES:
        e = Engine(device=0); e.set_terrain(ys, K.TERRAIN_STEP)
        for k_, v_ in opts.items():
            e.set_option(k_, v_)
        e.upload(pop)
        ms = []
        for _ in range(3):
            e.run_episodes(10000); ms.append(e.last_step_ms())
        res[mode] = (min(ms), e.ticks().sum())
        e.close()
    best = min(res, key=lambda m: res[m][0])
    print("pop %6d: " % n + "  ".join("%s %7.1f ms" % (m, res[m][0]) for m, _ in MODES) + "   best %s, %.3g creature-steps/s" % (
        best, res[best][1] / res[best][0] * 1e3), flush=True)
