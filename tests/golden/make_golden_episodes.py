"""Whole-episode fixtures produced by the UNMODIFIED reference ``evaluate()`` (REM2D_main.py:350-378) running on the
oracle-backed Box2D shim (oracle_box2d.py). Build container only.

    python tests/golden/make_golden_episodes.py

Output (committed): episodes_<enc>.npz = the flattened tables of N seeded random individuals as the reference's own
``Modular2D.reset`` built them (incl. controller parameters), and per individual the fitness ``evaluate`` returned and the
number of ``env.step`` calls it made. tests/test_reference_episodes.py asserts that ``rem2d_evaluate`` (oracle on CPU, CUDA
on the GPU box) reproduces both exactly.
"""
import os
import random
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
warnings.simplefilter("ignore")
import oracle_box2d  # noqa: E402
import ref_shim  # noqa: E402
from make_golden import record  # noqa: E402  (table of the world the reference just built)

N = {"direct": 200, "lsystem": 200, "ce": 120}


def main():
    sys.argv = sys.argv[:1]                      # REM2D_main.getEnv() parses the command line
    r2d = oracle_box2d.install()
    for enc, n in N.items():
        recs, fits, steps, seeds = [], [], [], []
        t0 = time.time()
        for i in range(n):
            seed = 104729 * (i + 1) + len(enc)
            random.seed(seed)
            ind = r2d.Individual.random(encoding=enc)
            if i % 5 == 4:                       # some mutated genomes as well
                for _ in range(2):
                    ind.genome.mutate(0.3, 0.3, 0.2)
            env = r2d.getEnv()
            calls = [0]
            orig = env.step

            def counting_step(action, _orig=orig, _calls=calls):
                _calls[0] += 1
                return _orig(action)
            env.step = counting_step
            fitness = r2d.evaluate(ind, EVALUATION_STEPS=10000, TREE_DEPTH=ind.tree_depth)
            del env.step
            # the table as the reference built it: reset again on the recording world (no stepping) and record
            recs.append(record(env, ind))
            # record() resets the env: controllers are fresh copies of the genome's (i_state as at the start of evaluate)
            fits.append(float(fitness)); steps.append(calls[0]); seeds.append(seed)
        nb = np.array([len(r["shape"]) for r in recs], np.int32)
        out = dict(seeds=np.array(seeds, np.int64), fitness=np.array(fits, np.float64), steps=np.array(steps, np.int32),
                   body_off=np.concatenate([[0], np.cumsum(nb)]).astype(np.int32))
        for k, dt in (("shape", np.uint8), ("hx", np.float32), ("hy", np.float32), ("x0", np.float32), ("y0", np.float32),
                      ("a0", np.float32), ("joint_parent", np.int16), ("lower", np.float32), ("upper", np.float32),
                      ("max_torque", np.float32), ("node_index", np.int32), ("type_ref", np.int16)):
            out[k] = np.array([v for r in recs for v in r[k]], dt)
        out["anchor_a"] = np.array([v for r in recs for v in r["anchor_a"]], np.float32).reshape(-1, 2)
        out["anchor_b"] = np.array([v for r in recs for v in r["anchor_b"]], np.float32).reshape(-1, 2)
        out["ctrl"] = np.array([v for r in recs for v in r["ctrl"]], np.float64).reshape(-1, 5)
        np.savez_compressed(os.path.join(HERE, "episodes_%s.npz" % enc), **out)
        print(enc, "individuals", n, "mean steps %.1f" % np.mean(steps), "max", max(steps), "fitness mean %.3f max %.3f" % (
            np.mean(fits), max(fits)), "%.1f s" % (time.time() - t0), "oracle worlds", oracle_box2d.OracleWorld.n_worlds_stepped)


if __name__ == "__main__":
    main()
