"""Debug helper (GPU box): step oracle and CUDA in lockstep, report the first creature/tick/field that differs."""
import random
import sys

import numpy as np

sys.path.insert(0, ".")
from gym_rem2d_b200 import Individual, constants as K, terrain
from gym_rem2d_b200.capi import Engine
from gym_rem2d_b200.flatten import flatten_population
from oracle.oracle import OracleEngine

enc, flat, n, seed, ticks = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
random.seed(seed)
pop = flatten_population([Individual.random(encoding=enc) for _ in range(n)])
xs, ys = terrain.flat_terrain() if flat else terrain.generate_terrain()
g, o = Engine(device=0), OracleEngine(threads=8)
for e in (g, o):
    e.set_terrain(ys, K.TERRAIN_STEP)
    e.upload(pop)
boff = pop.body_off
joff = pop.joint_off()
found = 0
bad = set()
for t in range(1, ticks + 1):
    g.step(1); o.step(1)
    sg, so = g.read_state(max_pairs=24), o.read_state(max_pairs=24)
    for c in range(pop.n_creatures):
        if c in bad:
            continue
        bs, js = slice(boff[c], boff[c + 1]), slice(joff[c], joff[c + 1])
        diffs = []
        for k in ("pose", "vel"):
            if not np.array_equal(sg[k][bs], so[k][bs]): diffs.append(k)
        for k in ("joint_impulse", "motor_speed", "limit_state"):
            if not np.array_equal(sg[k][js], so[k][js]): diffs.append(k)
        for k in ("alive", "ticks", "awake", "n_contacts", "n_touching", "touching_pairs", "touching_impulse", "wod"):
            if not np.array_equal(sg[k][c], so[k][c]): diffs.append(k)
        if diffs:
            bad.add(c)
            found += 1
            nb = boff[c + 1] - boff[c]
            print("tick %d creature %d nb=%d differs in %s" % (t, c, nb, diffs))
            print(" shapes", pop.shape[bs].tolist(), "parents", pop.joint_parent[js].tolist())
            for k in ("pose", "vel"):
                d = sg[k][bs].astype(np.float64) - so[k][bs]
                print(" ", k, "max abs diff", np.abs(d).max(), "bodies", np.nonzero(np.abs(d).max(axis=1))[0].tolist())
            print("  limit g", sg["limit_state"][js].tolist(), "o", so["limit_state"][js].tolist())
            print("  jimp diff", np.abs(sg["joint_impulse"][js].astype(np.float64) - so["joint_impulse"][js]).max() if nb > 1 else 0)
            print("  touching g", sg["n_touching"][c], sg["touching_pairs"][c][:sg["n_touching"][c]].tolist())
            print("  touching o", so["n_touching"][c], so["touching_pairs"][c][:so["n_touching"][c]].tolist())
            print("  timp g", sg["touching_impulse"][c][:sg["n_touching"][c]].tolist())
            print("  timp o", so["touching_impulse"][c][:so["n_touching"][c]].tolist())
            print("  pose g", sg["pose"][bs].tolist())
            print("  pose o", so["pose"][bs].tolist())
            print("  vel g", sg["vel"][bs].tolist())
            print("  vel o", so["vel"][bs].tolist())
            if found >= 3:
                print("counters g", g.counters()); print("counters o", o.counters())
                sys.exit(0)
print("no further divergence in", ticks, "ticks; diverged creatures:", sorted(bad))
print("counters g", g.counters()); print("counters o", o.counters())
