"""GPU diagnostic: lifetime distribution, fixed-horizon throughput per size class."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from gym_rem2d_b200 import constants as K, terrain
from gym_rem2d_b200.capi import Engine
from gym_rem2d_b200.population import random_population

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
pop = random_population(n, ("lsystem",), seed=2, cache_dir="/tmp/rem2d_cache")
xs, ys = terrain.generate_terrain()
e = Engine(device=0)
e.set_terrain(ys, K.TERRAIN_STEP)
fit, ticks = e.evaluate(pop, 10000)
fit, ticks = e.evaluate(pop, 10000)
print("episode ms", e.last_step_ms(), "steps", ticks.sum(), "rate", ticks.sum() / e.last_step_ms() * 1e3)
print("ticks pct", np.percentile(ticks, [0, 50, 90, 99, 99.9, 100]).tolist())
import ctypes
buf = (ctypes.c_float * 80)()
e.lib.rem2d_debug_class_timeline.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
ncls = e.lib.rem2d_debug_class_timeline(e.h, buf)
for k in range(ncls):
    if buf[5 * k + 1] > 0:
        print("class NB=%2d members=%6d warps=%4d begin %8.1f ms end %8.1f ms" % tuple(buf[5 * k + i] for i in range(5)))
if len(sys.argv) > 2 and sys.argv[2] == "timeline":
    sys.exit(0)
nb = np.diff(pop.body_off)
print("nb hist", np.bincount(nb).tolist())
for lo, hi in ((1, 2), (3, 4), (5, 8), (9, 12), (13, 16), (17, 22)):
    m = (nb >= lo) & (nb <= hi)
    print("class nb %d-%d: n=%d mean ticks %.1f max %d" % (lo, hi, m.sum(), ticks[m].mean() if m.any() else 0, ticks[m].max() if m.any() else 0))
# fixed horizon, no termination: pure per-tick throughput
e2 = Engine(device=0, terminate=0)
e2.set_terrain(ys, K.TERRAIN_STEP)
for lo, hi in ((1, 2), (3, 4), (5, 8), (9, 12), (13, 16), (17, 22), (1, 22)):
    idx = np.nonzero((nb >= lo) & (nb <= hi))[0]
    if len(idx) == 0:
        continue
    sub = pop.select(idx)
    e2.upload(sub)
    e2.step(100)
    e2.reset()
    e2.step(100)
    ms = e2.last_step_ms()
    c = e2.counters()
    print("fixed 100 ticks nb %d-%d: n=%d batches=%d ms=%.2f rate=%.3g c-steps/s  per-batch-tick=%.3f ms  jv/tick/creature=%.1f touching/tick=%.2f pos-iters=%.1f toi_events/tick=%.3f" % (
        lo, hi, len(idx), (len(idx) + 31) // 32, ms, len(idx) * 100 / ms * 1e3, ms / 100,
        c["joint_vsolves"] / 180 / c["ticks"], (c["p1_vsolves"] + c["m2_vsolves"]) / 180 / c["ticks"],
        c["joint_psolves"] / max(1, c["joint_vsolves"] / 180), c["toi_events"] / c["ticks"]))
