"""Parity of the oracle against REAL pybox2d (Box2D==2.3.10, /root/reference/requirements.txt:1) — runs wherever that wheel
is importable and is skipped elsewhere (it is not installable in the build container or on the GPU box: no wheel, no swig,
no network). Until this test has run green somewhere, every result of this repository is parity with
a RESTATEMENT of Box2D 2.3, not with Box2D itself ("parity unpinned").

The scene is built with the same pybox2d calls the reference makes (Modular2DEnv.py:294-306 terrain edges;
simple_module.py:286-298 / circular_module.py:191-202 bodies; module_utility.py:19-32 joints) from the committed flattened
tables, stepped with world.Step(1/50, 180, 60) under the reference's controller / P-control rule, and compared with the
oracle tick by tick: poses within 1e-4 relative over the first 100 ticks (BASELINE north_star), identical touching (body,
edge) pairs until the first divergence of more than the tolerance, and the episode fitness distribution.
"""
import math
import os

import numpy as np
import pytest

Box2D = pytest.importorskip("Box2D")

from gym_rem2d_b200 import constants as K, terrain  # noqa: E402
from gym_rem2d_b200.flatten import PopulationTable  # noqa: E402
from oracle.oracle import OracleEngine  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIELDS = ("body_off", "shape", "hx", "hy", "x0", "y0", "a0", "node_index", "type_ref", "joint_parent", "anchor_a", "anchor_b",
          "lower", "upper", "max_torque", "ctrl")


def build_world(pop, c, ys):
    from Box2D.b2 import circleShape, edgeShape, fixtureDef, polygonShape, revoluteJointDef
    world = Box2D.b2World()
    fd_edge = fixtureDef(shape=edgeShape(vertices=[(0, 0), (1, 1)]), friction=2.5, categoryBits=0x0001)
    for i in range(len(ys) - 1):
        fd_edge.shape.vertices = [(i * K.TERRAIN_STEP, ys[i]), ((i + 1) * K.TERRAIN_STEP, ys[i + 1])]
        world.CreateStaticBody(fixtures=fd_edge)
    b0, b1 = pop.body_off[c], pop.body_off[c + 1]
    bodies = []
    for b in range(b0, b1):
        if pop.shape[b] == 1:
            fd = fixtureDef(shape=circleShape(radius=float(pop.hx[b])), density=1, friction=0.1, restitution=0.0,
                            categoryBits=0x0020, maskBits=0x001)
        else:
            fd = fixtureDef(shape=polygonShape(box=(float(pop.hx[b]), float(pop.hy[b]))), density=1, friction=0.1, restitution=0.0,
                            categoryBits=0x0020, maskBits=0x001)
        bodies.append(world.CreateDynamicBody(position=(float(pop.x0[b]), float(pop.y0[b])), angle=float(pop.a0[b]), fixtures=fd))
    joints = []
    j0 = b0 - c
    for k in range(b1 - b0 - 1):
        j = j0 + k
        joints.append(world.CreateJoint(revoluteJointDef(
            bodyA=bodies[int(pop.joint_parent[j])], bodyB=bodies[k + 1],
            localAnchorA=tuple(float(v) for v in pop.anchor_a[j]), localAnchorB=tuple(float(v) for v in pop.anchor_b[j]),
            enableMotor=True, enableLimit=True, maxMotorTorque=float(pop.max_torque[j]), motorSpeed=0.0,
            lowerAngle=float(pop.lower[j]), upperAngle=float(pop.upper[j]))))
    return world, bodies, joints


def reference_ticks(pop, c, ys, n_ticks):
    """Poses after every tick under the reference's step() rule (Modular2DEnv.py:607-649), no termination."""
    world, bodies, joints = build_world(pop, c, ys)
    b0 = pop.body_off[c]
    ctrl = pop.ctrl[b0:b0 + len(bodies)].copy()
    out = []
    for _ in range(n_ticks):
        ctrl[:, 4] += ctrl[:, 2]
        outv = ctrl[:, 0] * np.sin(ctrl[:, 4] + ctrl[:, 1]) + ctrl[:, 3]
        for k, j in enumerate(joints):
            j.motorSpeed = (outv[k + 1] - j.angle) * 1.9
        world.Step(1.0 / 50, 180, 60)
        out.append([(b.position[0], b.position[1], b.angle) for b in bodies])
    return np.array(out, np.float64)


@pytest.mark.parametrize("enc", ["direct", "lsystem", "ce"])
def test_first_100_ticks_within_1e4_of_pybox2d(enc):
    z = np.load(os.path.join(GOLDEN, "episodes_%s.npz" % enc))
    pop = PopulationTable(*(z[k] for k in FIELDS))
    xs, ys = terrain.generate_terrain()
    o = OracleEngine(terminate=0, sincos_mode=1)           # libm sinf/cosf like upstream b2Rot::Set
    o.set_terrain(ys, K.TERRAIN_STEP)
    n = min(pop.n_creatures, 40)
    sub = pop.select(np.arange(n))
    o.upload(sub)
    ref = [reference_ticks(sub, c, ys, 100) for c in range(n)]
    worst = 0.0
    for t in range(100):
        o.step(1)
        pose = o.read_state()["pose"].astype(np.float64)
        for c in range(n):
            b0, b1 = sub.body_off[c], sub.body_off[c + 1]
            err = np.abs(pose[b0:b1] - ref[c][t]) / np.maximum(1.0, np.abs(ref[c][t]))
            worst = max(worst, float(err.max()))
    assert worst <= 1e-4, "oracle deviates from pybox2d by %g (relative) within the first 100 ticks" % worst
