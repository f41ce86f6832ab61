"""GPU diagnostic: steady-state tick latency of each class when run alone vs together (REM2D_TRACE samples)."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, ".")
from gym_rem2d_b200 import constants as K, terrain
from gym_rem2d_b200.capi import Engine
from gym_rem2d_b200.population import random_population

pop = random_population(65536, ("lsystem",), seed=2, cache_dir="/tmp/rem2d_cache")
nb = np.diff(np.asarray(pop.body_off))
xs, ys = terrain.generate_terrain()
os.environ["REM2D_TRACE"] = "1"
S = 1024
buf = np.zeros(1024 * S * 2, np.uint32)


def run(tag, mask):
    sub = pop.select(np.nonzero(mask)[0])
    e = Engine(device=0)
    e.set_terrain(ys, K.TERRAIN_STEP)
    e.upload(sub)
    e.run_episodes(10000)
    e.run_episodes(10000)
    e.lib.rem2d_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64]
    line = "%-28s n=%6d  %5.0f ms |" % (tag, mask.sum(), e.last_step_ms())
    for k in range(9):
        w = e.lib.rem2d_debug_trace(e.h, k, buf.ctypes.data_as(ctypes.c_void_p), buf.size)
        if w <= 0:
            continue
        a = buf[: w * S * 2].reshape(w, S, 2)
        t = a[:, :, 0].astype(np.int64)
        live = a[:, :, 1] & 0xff
        lat = []
        for i in range(w):
            v = t[i] > 0
            tv = t[i][v]
            if len(tv) < 40:
                continue
            full = (live[i][v] >= 28)[1:]
            d = np.diff(tv) / 4000.0
            d = d[10:][full[10:]]             # skip the fall phase, keep ticks with (almost) all lanes live
            if len(d):
                lat.append(np.median(d))
        line += " c%d: %dw %.2f ms" % (k, w, np.median(lat) if lat else float("nan"))
    print(line, flush=True)
    e.close()


sel = os.environ.get("MIX", "2..8")
if sel == "2..8":
    run("nb 2..8", (nb >= 2) & (nb <= 8))
elif sel == "all":
    run("all", nb >= 1)
elif sel == "2":
    run("nb 2 only", nb == 2)
