"""GPU experiment: whole-population episode time under different park policies / library variants (one gpurun call).
usage: python tools/sweep_policy.py [n_creatures]"""
import ctypes, os, sys, time
import numpy as np
sys.path.insert(0, ".")
from gym_rem2d_b200 import constants as K, terrain
from gym_rem2d_b200.capi import Engine
from gym_rem2d_b200.population import random_population

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
pop = random_population(n, ("lsystem",), seed=2, cache_dir="/tmp/rem2d_cache")
xs, ys = terrain.generate_terrain()
KNOBS = ("REM2D_PARK_LEAD_FROM", "REM2D_PARK_LEAD", "REM2D_SMALL_WEIGHT", "REM2D_SMEM_BUDGET_KB", "REM2D_PARK_TICKS", "REM2D_PARK_LATE", "REM2D_PARK_LATE_FROM", "REM2D_PARK_CAP", "REM2D_DRAIN_LANES")


def run(tag, lib=None, **env):
    for k in KNOBS:
        os.environ.pop(k, None)
    for k, v in env.items():
        os.environ["REM2D_" + k] = str(v)
    e = Engine(lib_path=lib, device=0)
    e.set_terrain(ys, K.TERRAIN_STEP)
    e.upload(pop)
    ms = []
    for _ in range(3):
        e.run_episodes(10000)
        ms.append(e.last_step_ms())
    ticks = e.ticks()
    buf = (ctypes.c_float * 80)()
    e.lib.rem2d_debug_class_timeline.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    ncls = e.lib.rem2d_debug_class_timeline(e.h, buf)
    ends = " ".join("%d:%.0f" % (buf[5 * k], buf[5 * k + 4]) for k in range(ncls) if buf[5 * k + 1] > 0)
    print("%-44s ms %s  rate %.3g  launches %d | class ends %s" % (tag, " ".join("%.0f" % m for m in ms), ticks.sum() / min(ms) * 1e3,
                                                                 e.launch_count(), ends), flush=True)
    e.close()
    return ticks


import json
cfgs = json.loads(os.environ.get("SWEEP_CFGS", "null")) or [
    ["default", None, {}],
    ["cap 1", None, {"PARK_CAP": 1.0}],
    ["park 224, cap 1", None, {"PARK_TICKS": 224, "PARK_CAP": 1.0}],
    ["park 256, cap 1", None, {"PARK_TICKS": 256, "PARK_CAP": 1.0}],
    ["park 320, cap 1", None, {"PARK_TICKS": 320, "PARK_CAP": 1.0}],
    ["no parking", None, {"PARK_TICKS": 0}],
]
base = None
for rnd in range(int(os.environ.get("SWEEP_ROUNDS", "2"))):
    for tag, lib, env in cfgs:
        t = run(tag, lib=lib, **env)
        if base is None:
            base = t
        assert np.array_equal(t, base), tag
print("all variants returned identical tick counts")
