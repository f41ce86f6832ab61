"""Lanes-per-creature sweep on EA-CONFIGURED populations (max_size 40, max_depth 7: creatures of up to 41 bodies, which the bench
population - LSystem defaults, <= 21 bodies - never has): the initial random population of run_deap and its tournament winners.
usage: python tools/ea_pop_sweep.py N "opt=value;..." ...      ("" = defaults)"""
import os, random, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_rem2d_b200 import constants as K, ea, terrain
from gym_rem2d_b200.population import concat

if __name__ == "__main__":
    n = int(sys.argv[1])
    cfg = ea.default_config(enc="lsystem")
    cfg["ea"]["batch_size"] = str(n)
    random.seed(2); np.random.seed(2)
    run = ea.run2D(cfg, "", workers=max(2, (os.cpu_count() or 4) - 2))          # pool before CUDA
    sizes = [len(c) for c in run._chunks(list(range(n)))]
    parts = run.pool.map(ea._random_chunk_packed, [(m, run.moduleList, cfg, random.getrandbits(48), run.TREE_DEPTH) for m in sizes])
    run.close()
    pop = concat([t for _, t in parts])
    from gym_rem2d_b200.capi import Engine
    xs, ys = terrain.generate_terrain()
    bounds = [1, 2, 4, 8, 12, 16, 22, 32, 44]
    rng = np.random.RandomState(0)
    parent_ticks = None
    for gen in range(3):
        nb = np.diff(pop.body_off)
        hist = [int(((nb > (bounds[i - 1] if i else 0)) & (nb <= b)).sum()) for i, b in enumerate(bounds)]
        print("generation %d: %d creatures, mean bodies %.1f, class histogram %s" % (gen, pop.n_creatures, nb.mean(), hist), flush=True)
        os.makedirs("/tmp/rem2d_cache", exist_ok=True)
        np.savez("/tmp/rem2d_cache/ea_pop_gen%d.npz" % gen, **{k: getattr(pop, k) for k in (
            "body_off", "shape", "hx", "hy", "x0", "y0", "a0", "node_index", "type_ref", "joint_parent", "anchor_a", "anchor_b", "lower",
            "upper", "max_torque", "ctrl")})
        ref = None
        for c in sys.argv[2:] or [""]:
            g = Engine(device=0); g.set_terrain(ys, K.TERRAIN_STEP)
            for kv in [kv for kv in c.split(";") if kv and kv != "PRIO"]:
                k_, v_ = kv.split("="); g.set_option(k_, float(v_))
            ms = []
            for _ in range(2):
                if "PRIO" in c.split(";") and parent_ticks is not None:
                    g.set_priority(parent_ticks)            # lifetimes of the (unmutated) parents: what ea.run2D passes
                f, t = g.evaluate(pop, K.EVALUATION_STEPS)
                ms.append(g.last_step_ms())
            if ref is None:
                ref = (f, t)
            print("   %-60s %s ms  %.3g creature-steps/s  identical %s" % (c or "(defaults)", " ".join("%.0f" % m for m in ms),
                  t.sum() / min(ms) * 1e3, bool(np.array_equal(f, ref[0]) and np.array_equal(t, ref[1]))), flush=True)
            g.close()
        fit, ticks = ref
        asp = rng.randint(0, len(fit), size=(len(fit), 4))
        win = asp[np.arange(len(fit)), np.argmax(fit[asp], axis=1)]
        pop = pop.select(win)
        parent_ticks = ticks[win].astype(np.float32)
