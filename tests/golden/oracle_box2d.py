"""A ``Box2D``-named module whose b2World is backed by the CPU oracle (SURVEY.md 8b "optional lower boundary", Appendix D).

With it the UNMODIFIED reference — ``REM2D_main.evaluate`` (REM2D_main.py:350-378) on top of ``Modular2D.reset/step``
(Modular2DEnv.py:565-653), the module ``create`` methods and ``create_joint`` — runs whole episodes in this container:
every pybox2d call it makes (call sites Modular2DEnv.py:144,226,301,572,634, simple_module.py:286-298,
circular_module.py:191-202, module_utility.py:19-32) lands here, world construction is recorded exactly as the recording
fake of ref_shim.py does, and ``world.Step`` advances a one-creature oracle world. Everything AROUND the physics — the
controllers, the P-controller, the wall of death, the reward and termination rules, the fitness latch and its step
accounting — is then executed by the reference's own Python, which pins those parts of the oracle / CUDA episode
semantics (rows a1, a2, a9, a10, a12) to reference-executed outputs. The physics inside Step is the oracle's restatement
on both sides, so this is NOT a pin of the Box2D arithmetic itself (that needs a real pybox2d: tests/test_pybox2d_parity.py).

Build container only (reads /root/reference); used by make_golden_episodes.py to write tests/golden/episodes_*.npz.
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ref_shim  # noqa: E402
from gym_rem2d_b200.flatten import PopulationTable  # noqa: E402
from oracle.oracle import OracleEngine  # noqa: E402


def table_from_world(w):
    """Flattened one-creature table from the recorded bodies / joints (as tests/golden/make_golden.py records them)."""
    bodies = w.dynamic_bodies
    slot = {id(b): i for i, b in enumerate(bodies)}
    nb = len(bodies)
    shape = np.array([1 if b.shape_kind == "circle" else 0 for b in bodies], np.uint8)
    hx = np.array([b.radius if b.shape_kind == "circle" else b.box[0] for b in bodies], np.float32)
    hy = np.array([0.0 if b.shape_kind == "circle" else b.box[1] for b in bodies], np.float32)
    x0 = np.array([b.position.x for b in bodies], np.float32)
    y0 = np.array([b.position.y for b in bodies], np.float32)
    a0 = np.array([b.angle for b in bodies], np.float32)
    for k, j in enumerate(w.joints):
        assert slot[id(j.bodyB)] == k + 1, "joint k must drive body k+1 (creation order)"
    jp = np.array([slot[id(j.bodyA)] for j in w.joints], np.int16)
    aa = np.array([j.kw["localAnchorA"] for j in w.joints], np.float32).reshape(-1, 2)
    ab = np.array([j.kw["localAnchorB"] for j in w.joints], np.float32).reshape(-1, 2)
    lo = np.array([j.kw["lowerAngle"] for j in w.joints], np.float32)
    up = np.array([j.kw["upperAngle"] for j in w.joints], np.float32)
    mt = np.array([j.kw["maxMotorTorque"] for j in w.joints], np.float32)
    ctrl = np.zeros((nb, 5), np.float64)          # controllers run in the reference's Python in this mode
    return PopulationTable(np.array([0, nb], np.int32), shape, hx, hy, x0, y0, a0, np.arange(nb, dtype=np.int32),
                           np.zeros(nb, np.int16), jp, aa, ab, lo, up, mt, ctrl)


class OracleWorld(ref_shim.b2World):
    """Recording world whose Step() is the oracle's b2World::Step."""
    n_worlds_stepped = 0

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self._eng = None

    def _build(self, dt, vel_it, pos_it):
        edges = [b.vertices for b in self.static_bodies]
        ys = np.array([e[0][1] for e in edges] + [edges[-1][1][1]], np.float64)
        step = edges[0][1][0] - edges[0][0][0]
        assert abs(step - 14.0 / 30.0) < 1e-12 and len(ys) == 200
        assert all(b.fixture_kw["friction"] == 2.5 for b in self.static_bodies)
        self._eng = OracleEngine(dt=dt, velocity_iterations=vel_it, position_iterations=pos_it)
        self._eng.set_terrain(ys, step)
        self._eng.upload(table_from_world(self))
        self._eng.lib.rem2d_oracle_world_step.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        OracleWorld.n_worlds_stepped += 1

    def Step(self, dt, vel_it, pos_it):
        super().Step(dt, vel_it, pos_it)
        if not self.dynamic_bodies:
            return
        if self._eng is None:
            self._build(dt, vel_it, pos_it)
        speeds = np.array([j.motorSpeed for j in self.joints], np.float32)      # SWIG setter: float32
        rc = self._eng.lib.rem2d_oracle_world_step(self._eng.h, 0, speeds.ctypes.data, len(speeds))
        assert rc == 0
        pose = self._eng.read_state()["pose"]
        for b, (x, y, a) in zip(self.dynamic_bodies, pose):
            b.position = ref_shim._Vec2(x, y)
            b.angle = float(a)


def install():
    """The reference's modules with Box2D.b2World = OracleWorld. Returns REM2D_main."""
    r2d = ref_shim.install()
    sys.modules["Box2D"].b2World = OracleWorld
    return r2d
