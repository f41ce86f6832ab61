"""GPU experiment: whole-population evaluation time of small populations, bulk (lane per creature) vs warp-per-creature mode."""
import os, sys
import numpy as np
sys.path.insert(0, ".")
from gym_rem2d_b200 import constants as K, terrain
from gym_rem2d_b200.capi import Engine
from gym_rem2d_b200.population import random_population

big = random_population(16384, ("lsystem",), seed=2, cache_dir="/tmp/rem2d_cache")
xs, ys = terrain.generate_terrain()
for n in (128, 512, 1024, 2048, 3552, 6144, 8192, 16384):
    pop = big.select(np.arange(n))
    res = {}
    for mode, mx in (("bulk", "0"), ("warp", "100000000")):
        os.environ["REM2D_WARP_MODE_MAX"] = mx
        e = Engine(device=0); e.set_terrain(ys, K.TERRAIN_STEP); e.upload(pop)
        ms = []
        for _ in range(3):
            e.run_episodes(10000); ms.append(e.last_step_ms())
        res[mode] = (min(ms), e.ticks().sum())
        e.close()
    print("pop %6d: bulk %7.1f ms  warp-per-creature %7.1f ms   (%.2fx)  %d creature-steps, best %.3g steps/s" % (
        n, res["bulk"][0], res["warp"][0], res["bulk"][0] / res["warp"][0], res["bulk"][1], res["bulk"][1] / min(res["bulk"][0], res["warp"][0]) * 1e3), flush=True)
