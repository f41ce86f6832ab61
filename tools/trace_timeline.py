"""GPU diagnostic: macro timeline of one whole-population evaluation from the REM2D_TRACE samples of the episode kernels
(per class: resident warps, live lanes and tick latency over time; start delay of the warps = shared-memory packing)."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, ".")
from gym_rem2d_b200 import constants as K, terrain
from gym_rem2d_b200.capi import Engine
from gym_rem2d_b200.population import random_population

if len(sys.argv) > 1 and sys.argv[1].endswith(".npz"):        # a population table saved by tools/ea_pop_sweep.py
    from gym_rem2d_b200.flatten import PopulationTable
    z = np.load(sys.argv[1])
    pop = PopulationTable(*(z[k] for k in ("body_off", "shape", "hx", "hy", "x0", "y0", "a0", "node_index", "type_ref", "joint_parent",
                                           "anchor_a", "anchor_b", "lower", "upper", "max_torque", "ctrl")))
else:
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    pop = random_population(n, ("lsystem",), seed=2, cache_dir="/tmp/rem2d_cache")
xs, ys = terrain.generate_terrain()
e = Engine(device=0)
e.set_terrain(ys, K.TERRAIN_STEP)
e.upload(pop)
e.run_episodes(10000)
e.set_option("trace", 1)
e.run_episodes(10000)
print("traced run: %.0f ms, %d creature-steps" % (e.last_step_ms(), e.ticks().sum()))
S = 1024
e.lib.rem2d_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64]
buf = np.zeros(1024 * S * 2, np.uint32)
traces = {}
t0 = None
for k in range(9):
    w = e.lib.rem2d_debug_trace(e.h, k, buf.ctypes.data_as(ctypes.c_void_p), buf.size)
    if w <= 0:
        continue
    a = buf[: w * S * 2].reshape(w, S, 2).copy()
    traces[k] = a
    valid = a[:, :, 0] > 0
    m = a[:, :, 0][valid].min()
    t0 = m if t0 is None else min(t0, m)
out = {}
BIN = max(50, int(e.last_step_ms() / 16 / 50) * 50)
for k, a in traces.items():
    t = (a[:, :, 0].astype(np.int64) - int(t0)) / 1000.0       # ms
    live = (a[:, :, 1] & 0xff).astype(np.int64)
    smid = (a[:, :, 1] >> 24).astype(np.int64)
    valid = a[:, :, 0] > 0
    nw = a.shape[0]
    start = np.array([t[w][valid[w]].min() for w in range(nw)])
    end = np.array([t[w][valid[w]].max() for w in range(nw)])
    print("class %d: %d warps, %d distinct SMs; warp start ms: min %.1f median %.1f max %.1f; warp end ms: median %.0f max %.0f"
          % (k, nw, len(np.unique(smid[valid])), start.min(), np.median(start), start.max(), np.median(end), end.max()))
    rows = []
    for b0 in range(0, int(end.max()) + BIN, BIN):
        lanes = 0.0; lat = []; running = 0
        for w in range(nw):
            tv = t[w][valid[w]]; lv = live[w][valid[w]]
            sel = (tv >= b0) & (tv < b0 + BIN)
            if sel.any():
                running += 1
                lanes += lv[sel].mean()
                d = np.diff(tv)[sel[:-1]] / 4.0
                if len(d):
                    lat.append(np.median(d))
        rows.append((b0, running, lanes, float(np.median(lat)) if lat else 0.0))
    for b0, running, lanes, lat in rows:
        print("   t=%4d ms  warps running %4d  live lanes %7.0f (%.0f%% of %d lanes/creatures)  tick latency %.2f ms"
              % (b0, running, lanes, 100.0 * lanes / (nw * 32), nw * 32, lat))
np.savez_compressed("gpurun_out/trace.npz", **{"c%d" % k: v for k, v in traces.items()})
