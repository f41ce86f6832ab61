"""Workload for compute-sanitizer (racecheck / memcheck / synccheck) over every execution mode of the episode kernel.

  compute-sanitizer --tool racecheck python tools/sanitize.py [n_creatures] [max_ticks]

Modes: bulk warps with lane refill + parked creatures finished by tail launches (a small shared-memory budget and an early
park threshold make both happen with a few hundred creatures), one warp per creature, and the stepping kernel. Results are
compared with the CPU oracle so that a sanitizer-clean run is also a correct one."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_rem2d_b200 import constants as K, terrain  # noqa: E402
from gym_rem2d_b200.capi import Engine  # noqa: E402
from gym_rem2d_b200.population import random_population  # noqa: E402
from oracle.oracle import OracleEngine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 600
max_ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 160
pop = random_population(n, ("lsystem",), seed=5, workers=4)
xs, ys = terrain.generate_terrain()
o = OracleEngine(threads=os.cpu_count() or 8)
o.set_terrain(ys, K.TERRAIN_STEP)
fo, to = o.evaluate(pop, max_ticks)
for name, env in (("queue G=1 + refill + tail G=32", {"REM2D_WARP_MODE_MAX": "0", "REM2D_GROUP_SHIFT": "0", "REM2D_SMEM_BUDGET_KB": "8",
                                                      "REM2D_PARK_TICKS": "60", "REM2D_PARK_CAP": "0.25"}),
                  ("queue G=4 + refill + tail G=8", {"REM2D_WARP_MODE_MAX": "0", "REM2D_GROUP_SHIFT": "2", "REM2D_SMEM_BUDGET_KB": "8",
                                                     "REM2D_PARK_TICKS": "60", "REM2D_PARK_CAP": "0.25", "REM2D_TAIL_GROUP_SHIFT": "3"}),
                  ("automatic group width (under-filled GPU)", {"REM2D_WARP_MODE_MAX": "0"}),
                  ("warp-per-creature", {"REM2D_WARP_MODE_MAX": "1000000"})):
    os.environ.update(env)
    g = Engine(device=0)
    g.set_terrain(ys, K.TERRAIN_STEP)
    fg, tg = g.evaluate(pop, max_ticks)
    print(name, "creatures", n, "ticks", int(tg.sum()), "equal to oracle:", bool(np.array_equal(fg, fo) and np.array_equal(tg, to)), flush=True)
    g.close()
    for k in env:
        os.environ.pop(k)
g = Engine(device=0)
g.set_terrain(ys, K.TERRAIN_STEP)
sub = pop.select(np.arange(min(n, 128)))
g.upload(sub)
g.step(40)
print("step kernel ok", int(g.read_state()["ticks"].sum()), flush=True)
