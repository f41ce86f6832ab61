"""Tick latency of the large class vs number of resident warps per SM (fixed horizon, step kernel)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from gym_rem2d_b200 import constants as K, terrain
from gym_rem2d_b200.capi import Engine
from gym_rem2d_b200.population import random_population
pop = random_population(65536, ("lsystem",), seed=2, cache_dir="/tmp/rem2d_cache")
xs, ys = terrain.generate_terrain()
nb = np.diff(pop.body_off)
big = np.nonzero(nb >= 17)[0]
e = Engine(device=0, terminate=0)
e.set_terrain(ys, K.TERRAIN_STEP)
for batches in (74, 148, 296, 444):
    sub = pop.select(big[: batches * 32])
    e.upload(sub)
    e.step(60); e.reset(); e.step(60)
    print("class22 batches=%d (%.1f warps/SM): %.3f ms per warp-tick, %.3g creature-steps/s" % (
        batches, batches / 148.0, e.last_step_ms() / 60, batches * 32 * 60 / e.last_step_ms() * 1e3))
