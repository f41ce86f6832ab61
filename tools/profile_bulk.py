"""Workload for the ncu --set full capture of the dominant launch: the bulk launch of the largest class (NB=22) of the
bench population in the PRODUCTION configuration (parking on: the first episode_kernel launch of an evaluation is the queue-mode
launch of the largest class; tail launches follow). `nopark` as second argument disables parking (one launch per class)."""
import os, sys
import numpy as np
sys.path.insert(0, ".")
if len(sys.argv) > 2 and sys.argv[2] == "nopark":
    os.environ["REM2D_PARK_TICKS"] = "0"
from gym_rem2d_b200 import constants as K, terrain
from gym_rem2d_b200.capi import Engine
from gym_rem2d_b200.population import random_population
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
pop = random_population(n, ("lsystem",), seed=2, cache_dir="/tmp/rem2d_cache")
xs, ys = terrain.generate_terrain()
e = Engine(device=0); e.set_terrain(ys, K.TERRAIN_STEP); e.upload(pop)
for _ in range(2):
    e.run_episodes(10000)
    print("ms", e.last_step_ms(), "launches", e.launch_count(), "creature-steps", e.ticks().sum(), flush=True)
