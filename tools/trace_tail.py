"""GPU diagnostic: what happens to parked creatures (REM2D_TRACE): wait between parking and the tail warp's start,
tail tick latency, concurrency of tail warps over time. usage: python tools/trace_tail.py [PARK_TICKS ...]"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, ".")
from gym_rem2d_b200 import constants as K, terrain
from gym_rem2d_b200.capi import Engine
from gym_rem2d_b200.population import random_population

pop = random_population(65536, ("lsystem",), seed=2, cache_dir="/tmp/rem2d_cache")
xs, ys = terrain.generate_terrain()
for arg in sys.argv[1:] or ["256"]:
    pt, tgs = (arg.split(":") + ["5"])[:2]
    pt = int(pt)
    e = Engine(device=0)
    e.set_option("trace", 1); e.set_option("park_ticks", pt); e.set_option("tail_group_shift", int(tgs))
    if pt < 200:
        e.set_option("park_cap", 0.25)
    e.set_terrain(ys, K.TERRAIN_STEP)
    e.upload(pop)
    e.run_episodes(10000)
    e.run_episodes(10000)
    ms = e.last_step_ms()
    e.lib.rem2d_debug_tail_trace.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64]
    e.lib.rem2d_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64]
    buf = np.zeros(1 << 22, np.uint32)
    t0 = None
    for k in range(9):
        w = e.lib.rem2d_debug_trace(e.h, k, buf.ctypes.data_as(ctypes.c_void_p), buf.size)
        if w > 0:
            a = buf[: w * 2048].reshape(w, 1024, 2)
            v = a[:, :, 0][a[:, :, 0] > 0]
            t0 = int(v.min()) if t0 is None else min(t0, int(v.min()))
    print("park at %d ticks, tail group shift %s: run %.0f ms" % (pt, tgs, ms))
    allr = []
    for k in range(9):
        n = e.lib.rem2d_debug_tail_trace(e.h, k, buf.ctypes.data_as(ctypes.c_void_p), buf.size)
        if n <= 0:
            continue
        a = buf[: n * 4].reshape(n, 4).astype(np.int64)
        parked, start, end, tk = (a[:, 0] - t0) / 1e3, (a[:, 1] - t0) / 1e3, (a[:, 2] - t0) / 1e3, a[:, 3]
        lat = (end - start) / np.maximum(tk, 1)
        print("  class %d: parked %5d  park time ms [min %.0f med %.0f max %.0f]  wait for tail warp ms [med %.1f p90 %.1f max %.1f]  "
              "tail ticks [med %d max %d]  tail tick latency ms [med %.2f p90 %.2f]  last end %.0f ms"
              % (k, n, parked.min(), np.median(parked), parked.max(), np.median(start - parked), np.percentile(start - parked, 90),
                 (start - parked).max(), np.median(tk), tk.max(), np.median(lat), np.percentile(lat, 90), end.max()))
        allr.append(np.stack([start, end], 1))
        worst = np.argmax(end)
        print("     last finisher: parked at %.0f ms, started %.0f ms, %d tail ticks, ended %.0f ms (%.2f ms/tick)"
              % (parked[worst], start[worst], tk[worst], end[worst], lat[worst]))
    if allr:
        r = np.concatenate(allr)
        for b0 in range(0, int(r[:, 1].max()) + 50, 50):
            conc = ((r[:, 0] < b0 + 25) & (r[:, 1] > b0 + 25)).sum()
            print("     t=%4d ms: %5d tail warps running" % (b0 + 25, conc))
    e.close()
