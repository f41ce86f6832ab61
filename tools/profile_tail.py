"""Workload for an ncu capture of tail-mode launches: big creatures, parked early (REM2D_PARK_TICKS=64)."""
import os, sys
import numpy as np
sys.path.insert(0, ".")
os.environ["REM2D_PARK_TICKS"] = "64"
from gym_rem2d_b200 import constants as K, terrain
from gym_rem2d_b200.capi import Engine
from gym_rem2d_b200.population import random_population
pop = random_population(8192, ("lsystem",), seed=2, cache_dir="/tmp/rem2d_cache")
nb = np.diff(np.asarray(pop.body_off))
sub = pop.select(np.nonzero(nb >= 17)[0][:512])
xs, ys = terrain.generate_terrain()
e = Engine(device=0); e.set_terrain(ys, K.TERRAIN_STEP); e.upload(sub)
e.run_episodes(10000)
print("ms", e.last_step_ms(), "launches", e.launch_count(), "ticks", e.ticks().sum())
