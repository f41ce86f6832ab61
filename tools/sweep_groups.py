"""Run time of one evaluation of the bench population for different lanes-per-creature settings (REM2D_CLASS_GS)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_rem2d_b200 import constants as K, terrain  # noqa: E402
from gym_rem2d_b200.capi import Engine  # noqa: E402
from gym_rem2d_b200.population import random_population  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
configs = sys.argv[2:] or ["0,0,0,0,0,0,0,0,0", "0,0,1,2,2,2,2,3,3", "0,0,0,1,2,2,2,3,3", "0,0,1,1,1,1,1,2,2", "0,0,1,2,3,3,3,3,3", "0,0,0,2,2,2,3,3,3"]
pop = random_population(n, ("lsystem",), seed=2, workers=os.cpu_count(), cache_dir="/tmp/rem2d_cache")
xs, ys = terrain.generate_terrain()
ref = None
for cfg in configs:
    env = dict(kv.split("=") for kv in cfg.split(";")[1:]) if ";" in cfg else {}
    os.environ["REM2D_CLASS_GS"] = cfg.split(";")[0]
    for k_, v_ in env.items():
        os.environ[k_] = v_
    lib = env.pop("LIB", None)
    prio = env.pop("PRIO", None)
    g = Engine(device=0, lib_path=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gym_rem2d_b200", "csrc", lib) if lib else None)
    g.set_terrain(ys, K.TERRAIN_STEP)
    if prio and ref is not None:
        g.set_priority(ref[1].astype(np.float32))      # a perfect lifetime hint (what an EA approximates with the parents' lifetimes)
    g.upload(pop)
    ms = []
    for i in range(3):
        g.run_episodes(K.EVALUATION_STEPS)
        ms.append(g.last_step_ms())
    f, t = g.fitness(), g.ticks()
    if ref is None:
        ref = (f, t)
    same = bool(np.array_equal(f, ref[0]) and np.array_equal(t, ref[1]))
    print("class_gs %-22s %s ms %s  creature-steps/s %.3e  identical to first config: %s" % (
        cfg, " ".join("%.0f" % m for m in ms), "", t.sum() / (min(ms) * 1e-3), same), flush=True)
    g.close()
    for k_ in env:
        os.environ.pop(k_)
