"""Checkpoints are written under the REFERENCE's class paths and are readable by the reference's own
classes: REM2D_main.Individual / DataAnalysis.FitnessData and everything reachable from them (REM2D_main.py:311-329,
DataAnalysis.py:39-56, Experiments/Load_Best.py:7-37)."""
import os
import pickle
import pickletools
import random
import subprocess
import sys

import numpy as np
import pytest

from gym_rem2d_b200 import Individual, ea, refpickle
from gym_rem2d_b200.flatten import flatten_population

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference/ModularER_2D"


def population():
    random.seed(11)
    inds = [Individual.random(encoding=e) for e in ("direct", "lsystem", "ce") for _ in range(6)]
    for k, ind in enumerate(inds):
        ind.fitness = 1.0 + k
        if k % 2:
            Individual.mutate(0.3, 0.3, 0.2, ind)
    return inds


def test_pickles_name_only_reference_classes_and_round_trip():
    inds = population()
    data = refpickle.dumps(inds)
    globals_ = {arg for op, arg, _ in pickletools.genops(data) if op.name in ("GLOBAL", "STACK_GLOBAL") and arg}
    assert not [g for g in globals_ if "gym_rem2d_b200" in g], globals_
    assert {"REM2D_main Individual", "Encodings.lsystem LSystem", "Encodings.direct_encoding DirectEncoding",
            "Encodings.network_encoding NN_enc", "Tree Node" if False else "REM2D_main Individual"} <= globals_
    assert Individual.__module__ == "gym_rem2d_b200.individual"          # the switch is undone
    back = refpickle.loads(data)
    assert [type(b) for b in back] == [Individual] * len(inds)
    a, b = flatten_population(inds), flatten_population(back)
    for k in ("body_off", "shape", "hx", "hy", "x0", "y0", "a0", "joint_parent", "anchor_a", "anchor_b", "ctrl"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    fd = ea.FitnessData()
    fd.addFitnessData([1.0, 2.0, 4.0], 1)
    assert b"DataAnalysis" in refpickle.dumps(fd) and refpickle.loads(refpickle.dumps(fd)).p_100 == [4.0]
    # plain pickle of this package's objects still works (and is what refpickle.load also accepts)
    assert type(refpickle.loads(pickle.dumps(inds[0]))) is Individual


READER = r"""
import os, pickle, sys, warnings
import numpy as np
warnings.simplefilter("ignore")
sys.path.insert(0, os.path.join(sys.argv[1], "tests", "golden"))
import ref_shim
r2d = ref_shim.install()                      # the UNMODIFIED reference modules (fake Box2D / gym underneath)
from make_golden import record
import DataAnalysis
pop = pickle.load(open(sys.argv[2], "rb"))    # plain pickle.load, as Load_Best.py / run2D.run do
fd = pickle.load(open(sys.argv[3], "rb"))
assert type(fd) is DataAnalysis.FitnessData and len(fd.avg) == 1
assert all(type(p) is r2d.Individual for p in pop)
assert type(pop[0].genome).__module__.startswith("Encodings.")
env = ref_shim.reference_env()
recs = [record(env, ind) for ind in pop]      # reference genome.create + Modular2D.reset on the recording world
nb = np.array([len(r["shape"]) for r in recs])
out = {"nb": nb, "fitness": np.array([p.fitness for p in pop])}
for k in ("hx", "hy", "x0", "y0", "a0", "joint_parent"):
    out[k] = np.array([v for r in recs for v in r[k]], np.float64)
out["ctrl"] = np.array([v for r in recs for v in r["ctrl"]], np.float64)
np.savez(sys.argv[4], **out)
"""


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="needs the reference tree (build container only)")
def test_the_reference_unpickles_and_expands_our_checkpoints(tmp_path):
    inds = population()
    refpickle.dump(inds, tmp_path / "s_pop0")
    fd = ea.FitnessData()
    fd.addFitnessData([i.fitness for i in inds], 1)
    refpickle.dump(fd, tmp_path / "s_")
    script = tmp_path / "reader.py"
    script.write_text(READER)
    subprocess.check_call([sys.executable, str(script), ROOT, str(tmp_path / "s_pop0"), str(tmp_path / "s_"), str(tmp_path / "out.npz")])
    z = np.load(tmp_path / "out.npz")
    ours = flatten_population(inds)
    assert np.array_equal(z["nb"], np.diff(ours.body_off))
    assert np.array_equal(z["fitness"], [i.fitness for i in inds])
    for k in ("hx", "hy", "x0", "y0", "a0", "joint_parent"):
        assert np.array_equal(z[k], getattr(ours, k).astype(np.float64)), k
    assert np.array_equal(z["ctrl"], ours.ctrl)
