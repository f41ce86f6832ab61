"""Whole episodes of the UNMODIFIED reference ``evaluate()`` (REM2D_main.py:350-378 over Modular2D.reset/step,
Modular2DEnv.py:565-653) — run in the build container on the oracle-backed ``Box2D`` shim (tests/golden/oracle_box2d.py,
make_golden_episodes.py) — against ``rem2d_evaluate``: the fitness the reference returned and the number of env.step calls
it made must be reproduced exactly, by the oracle (CPU) and by the CUDA library (GPU box).

What this pins to reference-executed outputs: the episode semantics around world.Step — controller update order and
phase, the P-controller and its float32 motor-speed write, the wall of death, the reward / termination rules, the fitness
latch and the step accounting — over complete episodes with live physics, not
just 5 ticks on frozen bodies (control_pin.json). The Box2D arithmetic inside Step is the oracle's on both sides.
"""
import os

import numpy as np
import pytest

from gym_rem2d_b200 import constants as K, terrain
from gym_rem2d_b200.flatten import PopulationTable
from oracle.oracle import OracleEngine

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIELDS = ("body_off", "shape", "hx", "hy", "x0", "y0", "a0", "node_index", "type_ref", "joint_parent", "anchor_a", "anchor_b",
          "lower", "upper", "max_torque", "ctrl")


def load(enc):
    z = np.load(os.path.join(GOLDEN, "episodes_%s.npz" % enc))
    return PopulationTable(*(z[k] for k in FIELDS)), z["fitness"], z["steps"]


@pytest.mark.parametrize("enc", ["direct", "lsystem", "ce"])
def test_oracle_reproduces_the_reference_evaluate_loop(enc):
    pop, fit_ref, steps_ref = load(enc)
    xs, ys = terrain.generate_terrain()
    o = OracleEngine(threads=4)
    o.set_terrain(ys, K.TERRAIN_STEP)
    fit, ticks = o.evaluate(pop, K.EVALUATION_STEPS)
    assert np.array_equal(ticks, steps_ref), "env.step calls differ for %d individuals" % (ticks != steps_ref).sum()
    assert np.array_equal(fit, fit_ref), "fitness differs for %d individuals" % (fit != fit_ref).sum()
    assert steps_ref.min() >= 20 and fit_ref.max() > 5.0 and len(fit_ref) >= 100


@pytest.mark.gpu
@pytest.mark.parametrize("enc", ["direct", "lsystem", "ce"])
def test_cuda_reproduces_the_reference_evaluate_loop(enc, monkeypatch):
    from gym_rem2d_b200.capi import Engine
    pop, fit_ref, steps_ref = load(enc)
    xs, ys = terrain.generate_terrain()
    for mode in ("0", "1000000"):                    # queue mode with groups, and a warp per creature
        monkeypatch.setenv("REM2D_WARP_MODE_MAX", mode)
        g = Engine(device=0)
        g.set_terrain(ys, K.TERRAIN_STEP)
        fit, ticks = g.evaluate(pop, K.EVALUATION_STEPS)
        assert np.array_equal(ticks, steps_ref) and np.array_equal(fit, fit_ref), mode
