"""Where the time of a tick goes, per execution mode and lanes-per-creature: warp-cycles per tick phase from a diagnostic
build of the library (make -C gym_rem2d_b200/csrc B=build_pt OUT=librem2d_cuda_pt.so EXTRA=-DREM2D_PHASE_TIMING).
usage: python tools/phase_breakdown.py [n_creatures] [option=value ...]"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, ".")
from gym_rem2d_b200 import constants as K, terrain
from gym_rem2d_b200.capi import Engine
from gym_rem2d_b200.population import random_population

PHASES = ["loop/queue", "controllers", "collide", "stage", "schedule", "velocity", "store+integrate", "position", "finalize", "find_new",
          "toi_scan", "toi_events", "build_world", "finish/park", "-", "-"]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
opts = dict(a.split("=") for a in sys.argv[2:])
big = random_population(max(n, 32768) if n <= 32768 else n, ("lsystem",), seed=2, cache_dir="/tmp/rem2d_cache")
pop = big.select(np.arange(n)) if n < big.n_creatures else big
xs, ys = terrain.generate_terrain()
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gym_rem2d_b200", "csrc", "librem2d_cuda_pt.so")
e = Engine(device=0, lib_path=lib)
for k_, v_ in opts.items():
    e.set_option(k_, float(v_))
e.set_terrain(ys, K.TERRAIN_STEP)
e.upload(pop)
e.run_episodes(K.EVALUATION_STEPS)
e.run_episodes(K.EVALUATION_STEPS)
print("pop %d options %s: %.0f ms, %d creature-steps" % (n, opts, e.last_step_ms(), e.ticks().sum()))
out = np.zeros(12 * 16, np.uint64)
e.lib.rem2d_debug_phases.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
assert e.lib.rem2d_debug_phases(e.h, out.ctypes.data_as(ctypes.c_void_p)) == 0
out = out.reshape(2, 6, 16).astype(np.float64)
for mode in range(2):
    for gs in range(6):
        tot = out[mode, gs].sum()
        if tot == 0:
            continue
        print("%s launches, %2d lanes per creature: %.3g warp-cycles" % (("queue", "tail")[mode], 1 << gs, tot))
        for i in np.argsort(-out[mode, gs]):
            if out[mode, gs, i] > 0.002 * tot:
                print("      %-16s %5.1f %%" % (PHASES[i], 100 * out[mode, gs, i] / tot))
