"""Import shim that lets the UNMODIFIED reference Python (/root/reference/ModularER_2D) run in this
container, where Box2D / gym / matplotlib / neat / deap / termcolor are not installed.

Only used by tests/golden/make_golden.py (fixture generation, build container only). Nothing under
tests/ -m gpu, bench.py or smoke() imports this: /root/reference does not exist on the GPU box.

What is faked:
  * Box2D: a *recording* world. Bodies keep position/angle rounded to float32 exactly where pybox2d
    (SWIG float32 members) would round them; joints keep the def kwargs. Step() does nothing.
  * gym: Env, EzPickle, spaces.Box, utils.seeding.np_random (gym 0.18 algorithm), registration.register.
  * matplotlib: get_cmap returning a dummy colour function.
  * deap / termcolor: empty stubs so REM2D_main imports.
The reference assumes a case-insensitive filesystem (``from Encodings import lsystem``), so the
encoding files are loaded under the lower-case names it imports.
"""
import hashlib
import importlib.util
import os
import struct
import sys
import types

import numpy as np

REF_ROOT = "/root/reference/ModularER_2D"


def _f32(x):
    return float(np.float32(x))


# ----------------------------------------------------------------------------------------------
# fake Box2D
# ----------------------------------------------------------------------------------------------
class _Vec2:
    def __init__(self, x, y):
        self.x = _f32(x)
        self.y = _f32(y)

    def __getitem__(self, i):
        return (self.x, self.y)[i]

    def __iter__(self):
        return iter((self.x, self.y))

    def __len__(self):
        return 2


class _Bag:
    def __init__(self, **kw):
        self.__dict__.update(kw)
        self.kw = kw


class _Shape(_Bag):
    pass


def polygonShape(**kw):
    return _Shape(kind="polygon", **kw)


def edgeShape(**kw):
    return _Shape(kind="edge", **kw)


def circleShape(**kw):
    return _Shape(kind="circle", **kw)


def fixtureDef(**kw):
    return _Bag(**kw)


def revoluteJointDef(**kw):
    return _Bag(**kw)


class contactListener:
    def __init__(self):
        pass


class _Body:
    def __init__(self, kind, position=(0.0, 0.0), angle=0.0, fixtures=None):
        self.kind = kind
        self.position = _Vec2(position[0], position[1])
        self.angle = _f32(angle)
        fx = fixtures
        shape = fx.shape
        # snapshot of the shape (the reference mutates one shared fixtureDef for all terrain edges)
        self.shape_kind = shape.kind
        if shape.kind == "edge":
            self.vertices = [tuple(v) for v in shape.vertices]
        elif shape.kind == "polygon":
            if "box" in shape.kw:
                self.box = tuple(shape.box)
            else:
                self.vertices = [tuple(v) for v in shape.vertices]
        else:
            self.radius = shape.radius
        self.fixture_kw = {k: v for k, v in fx.kw.items() if k != "shape"}
        self.color1 = None
        self.color2 = None


class _Joint:
    def __init__(self, d):
        self.bodyA = d.bodyA
        self.bodyB = d.bodyB
        self.kw = dict(d.kw)
        self.motorSpeed = 0.0

    @property
    def angle(self):
        # b2RevoluteJoint::GetJointAngle, float32 arithmetic, referenceAngle = 0
        return float(np.float32(self.bodyB.angle) - np.float32(self.bodyA.angle))


class b2World:
    def __init__(self, *a, **kw):
        self.static_bodies = []
        self.dynamic_bodies = []
        self.joints = []
        self.n_steps = 0
        self.contactListener = None

    def CreateStaticBody(self, **kw):
        b = _Body("static", **kw)
        self.static_bodies.append(b)
        return b

    def CreateDynamicBody(self, **kw):
        b = _Body("dynamic", **kw)
        self.dynamic_bodies.append(b)
        return b

    def CreateJoint(self, d):
        j = _Joint(d)
        self.joints.append(j)
        return j

    def DestroyBody(self, b):
        pass

    def Step(self, dt, vel_it, pos_it):
        self.n_steps += 1
        self.last_step_args = (dt, vel_it, pos_it)


def b2CircleShape(**kw):
    return _Shape(kind="circle", **kw)


class b2RevoluteJoint:
    pass


# ----------------------------------------------------------------------------------------------
# gym 0.18 seeding (gym/utils/seeding.py): np_random(seed) -> RandomState seeded with the uint32
# words of sha512(str(seed))[:8]  (hash_seed + _bigint_from_bytes + _int_list_from_bigint)
# ----------------------------------------------------------------------------------------------
def _gym_np_random(seed=None):
    if seed is None:
        seed = int.from_bytes(os.urandom(8), "little")
    h = hashlib.sha512(str(seed).encode("utf8")).digest()[:8]
    # _bigint_from_bytes: little-endian uint32 words, accumulated little-endian
    padded = h + b"\0" * ((4 - len(h) % 4) % 4)
    n_words = len(padded) // 4
    words = struct.unpack("{}I".format(n_words), padded)
    big = 0
    for i, w in enumerate(words):
        big += 2 ** (32 * i) * w
    # _int_list_from_bigint
    ints = []
    while big > 0:
        big, mod = divmod(big, 2 ** 32)
        ints.append(mod)
    if not ints:
        ints = [0]
    rng = np.random.RandomState()
    rng.seed(ints)
    return rng, seed


class _Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low = np.asarray(low, dtype=dtype)
        self.high = np.asarray(high, dtype=dtype)
        self.dtype = dtype
        self.shape = self.low.shape

    def sample(self):
        return np.zeros(self.shape, dtype=self.dtype)


def install():
    """Register the stub modules and load the reference's encodings. Idempotent."""
    if "REM2D_main" in sys.modules:
        return sys.modules["REM2D_main"]
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)

    box2d = types.ModuleType("Box2D")
    b2 = types.ModuleType("Box2D.b2")
    for name, obj in dict(polygonShape=polygonShape, edgeShape=edgeShape, circleShape=circleShape,
                          fixtureDef=fixtureDef, revoluteJointDef=revoluteJointDef,
                          contactListener=contactListener).items():
        setattr(b2, name, obj)
        setattr(box2d, name, obj)
    box2d.b2 = b2
    box2d.b2World = b2World
    box2d.b2CircleShape = b2CircleShape
    box2d.b2RevoluteJoint = b2RevoluteJoint
    sys.modules["Box2D"] = box2d
    sys.modules["Box2D.b2"] = b2

    gym = types.ModuleType("gym")
    gym.__path__ = []

    class Env:
        pass

    gym.Env = Env
    spaces = types.ModuleType("gym.spaces")
    spaces.Box = _Box
    utils = types.ModuleType("gym.utils")
    utils.__path__ = []

    class EzPickle:
        def __init__(self, *a, **kw):
            pass

    utils.EzPickle = EzPickle
    utils.colorize = lambda s, *a, **k: s
    seeding = types.ModuleType("gym.utils.seeding")
    seeding.np_random = _gym_np_random
    utils.seeding = seeding
    envs = types.ModuleType("gym.envs")
    envs.__path__ = []
    registration = types.ModuleType("gym.envs.registration")
    registry = {}

    def register(id, entry_point=None, max_episode_steps=None, **kw):
        registry[id] = (entry_point, max_episode_steps)

    registration.register = register
    envs.registration = registration

    def make(env_id):
        entry, _ = registry[env_id]
        mod, cls = entry.split(":")
        m = importlib.import_module(mod)
        return getattr(m, cls)()

    gym.make = make
    gym.spaces = spaces
    gym.utils = utils
    gym.envs = envs
    sys.modules.update({"gym": gym, "gym.spaces": spaces, "gym.utils": utils,
                        "gym.utils.seeding": seeding, "gym.envs": envs,
                        "gym.envs.registration": registration})

    mpl = types.ModuleType("matplotlib")
    mpl.__path__ = []
    plt = types.ModuleType("matplotlib.pyplot")
    plt.get_cmap = lambda name=None: (lambda v: (0.5, 0.5, 0.5, 1.0))
    cm = types.ModuleType("matplotlib.cm")
    cm.viridis = lambda v: (0.5, 0.5, 0.5, 1.0)
    patches = types.ModuleType("matplotlib.patches")
    transforms = types.ModuleType("matplotlib.transforms")
    axes = types.ModuleType("matplotlib.axes")
    axes.__path__ = []
    _axes = types.ModuleType("matplotlib.axes._axes")

    class _Log:
        def setLevel(self, *a):
            pass

    _axes._log = _Log()
    axes._axes = _axes
    mpl.pyplot = plt
    mpl.cm = cm
    mpl.patches = patches
    mpl.transforms = transforms
    mpl.axes = axes
    sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt, "matplotlib.cm": cm,
                        "matplotlib.patches": patches, "matplotlib.transforms": transforms,
                        "matplotlib.axes": axes,
                        "matplotlib.axes._axes": _axes})

    for name in ("deap", "termcolor", "neat"):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules.setdefault(name, m)
    deap = sys.modules["deap"]
    for sub in ("base", "tools", "algorithms"):
        sm = types.ModuleType("deap." + sub)
        setattr(deap, sub, sm)
        sys.modules["deap." + sub] = sm
    sys.modules["termcolor"].colored = lambda s, *a, **k: s
    sys.modules["termcolor"].cprint = print

    class _DefaultGenome:  # enough for ``class CPPN_genome(neat.DefaultGenome)`` to be defined
        def __init__(self, key):
            self.key = key

    sys.modules["neat"].DefaultGenome = _DefaultGenome

    # the reference imports its encodings under lower-case module names
    import Encodings  # noqa: F401  (real package from the reference)
    for lower, fname in (("abstract_encoding", "Abstract_Encoding.py"),
                         ("cellular_encoding", "Cellular_Encoding.py"),
                         ("direct_encoding", "Direct_Encoding.py"),
                         ("lsystem", "LSystem.py"),
                         ("network_encoding", "Network_Encoding.py")):
        full = "Encodings." + lower
        spec = importlib.util.spec_from_file_location(full, os.path.join(REF_ROOT, "Encodings", fname))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[full] = mod
        setattr(sys.modules["Encodings"], lower, mod)
        spec.loader.exec_module(mod)

    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import REM2D_main
    return REM2D_main


def reference_env(flat=False):
    """A reference Modular2D env instance on the recording fake world."""
    install()
    import gym_rem2D.envs.Modular2DEnv as m2d
    m2d.MAX_PERTURBANCE_TERRAIN = 0 if flat else 24
    return m2d.Modular2D()
