"""What a fused-multiply-add build would buy and cost: librem2d_cuda_fma.so is the same source compiled
with -fmad=true (make -C gym_rem2d_b200/csrc B=build_fma OUT=librem2d_cuda_fma.so FMAD=true). It is NOT bit-identical to the
oracle any more, so it is judged by the north star's own tolerance (1e-4 relative over the first 100 ticks) and by the
fitness distribution of whole episodes; the default library stays the exact one."""
import os, sys
import numpy as np
sys.path.insert(0, ".")
from gym_rem2d_b200 import constants as K, terrain
from gym_rem2d_b200.capi import Engine
from gym_rem2d_b200.population import random_population

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FMA = os.path.join(HERE, "gym_rem2d_b200", "csrc", "librem2d_cuda_fma.so")


def ks(a, b):
    a, b = np.sort(a), np.sort(b)
    allv = np.concatenate([a, b])
    return float(np.abs(np.searchsorted(a, allv, side="right") / len(a) - np.searchsorted(b, allv, side="right") / len(b)).max())


xs, ys = terrain.generate_terrain()
# 1. trajectories: 2048 creatures, fixed horizon of 100 ticks
pop = random_population(2048, ("lsystem",), seed=2, cache_dir="/tmp/rem2d_cache")
poses = {}
for name, lib in (("exact", None), ("fma", FMA)):
    e = Engine(device=0, lib_path=lib, terminate=0)
    e.set_terrain(ys, K.TERRAIN_STEP)
    e.upload(pop)
    out = []
    for t in (10, 15, 25, 50):
        e.step(t)
        out.append(e.read_state()["pose"].astype(np.float64))
    poses[name] = out
    e.close()
for i, t in enumerate((10, 25, 50, 100)):
    a, b = poses["exact"][i], poses["fma"][i]
    rel = np.abs(a - b) / np.maximum(1.0, np.abs(a))
    per_creature = np.maximum.reduceat(rel.max(1), pop.body_off[:-1])
    print("after %3d ticks: max rel pose difference %.3g, median per creature %.3g, creatures within 1e-4: %.1f %%" % (
        t, rel.max(), np.median(per_creature), 100 * np.mean(per_creature <= 1e-4)))
# 2. whole episodes: fitness distribution and run time, bench population
big = random_population(65536, ("lsystem",), seed=2, cache_dir="/tmp/rem2d_cache")
res = {}
for name, lib in (("exact", None), ("fma", FMA)):
    e = Engine(device=0, lib_path=lib)
    e.set_terrain(ys, K.TERRAIN_STEP)
    e.upload(big)
    ms = []
    for _ in range(3):
        e.run_episodes(K.EVALUATION_STEPS)
        ms.append(e.last_step_ms())
    res[name] = (e.fitness(), e.ticks(), min(ms))
    e.close()
fe, te, me = res["exact"]
ff, tf, mf = res["fma"]
print("whole episodes, 65536 creatures: exact %.0f ms (%.3g steps/s), fma %.0f ms (%.3g steps/s): %.2fx" % (
    me, te.sum() / me * 1e3, mf, tf.sum() / mf * 1e3, me / mf))
print("fitness: KS %.4f, mean %.4f vs %.4f, identical lifetimes %.1f %%, identical fitness %.1f %%" % (
    ks(fe, ff), fe.mean(), ff.mean(), 100 * np.mean(te == tf), 100 * np.mean(fe == ff)))
