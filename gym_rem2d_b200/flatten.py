"""Genotype-expanded Tree -> flattened SoA body/joint/controller table (the input of librem2d).

Restates the robot assembly of the reference *without any physics objects*:
``Modular2D.create_robot`` / ``get_component_index`` / ``create_component`` (Modular2DEnv.py:326-333,
425-475,517-563), ``Standard2D.get_global_position_of_connection_site`` + ``create``
(simple_module.py:147-199,231-313), ``Circular2D.create`` (circular_module.py:157-221) and
``module_utility.create_joint`` (module_utility.py:7-33).

pybox2d stores body position/angle, shape extents, anchors, limits and torque as float32, and the
reference reads the *rounded* pose of the parent/child back into its double arithmetic for the next
level of the tree. ``_f32`` marks exactly those round trips, which is what makes
the emitted table bit-exact against the reference (tests/test_flatten.py, golden fixtures).
"""
import math
import struct
from dataclasses import dataclass, field

import numpy as np

from . import constants as K
from .modules import Circular2D
from .encodings.lsystem import LSystem


_F32 = struct.Struct("f")


def _f32(v):
    """``v`` rounded to float32 (C cast, round-to-nearest-even - what pybox2d's float members do), as a Python float."""
    try:
        return _F32.unpack(_F32.pack(v))[0]
    except OverflowError:                    # beyond the float32 range: inf, like the cast
        return float(np.float32(v))


_JOINT_LOWER32, _JOINT_UPPER32 = _f32(K.JOINT_LOWER), _f32(K.JOINT_UPPER)


@dataclass
class CreatureTable:
    """One creature. Bodies are in creation order; joint j connects ``joint_parent[j]`` -> body j+1."""
    shape: list = field(default_factory=list)        # 0 box, 1 circle
    hx: list = field(default_factory=list)           # half width  (circle: radius)
    hy: list = field(default_factory=list)           # half height (circle: 0)
    x0: list = field(default_factory=list)
    y0: list = field(default_factory=list)
    a0: list = field(default_factory=list)
    node_index: list = field(default_factory=list)
    type_ref: list = field(default_factory=list)
    joint_parent: list = field(default_factory=list)
    anchor_a: list = field(default_factory=list)     # (x, y) in the parent frame
    anchor_b: list = field(default_factory=list)     # (x, y) in the child frame
    lower: list = field(default_factory=list)
    upper: list = field(default_factory=list)
    max_torque: list = field(default_factory=list)
    ctrl: list = field(default_factory=list)         # per body: (amplitude, phase, frequency, offset, i_state)
    expressed: list = field(default_factory=list)    # per *node*: body slot or -1

    @property
    def n_bodies(self):
        return len(self.shape)


def _module_of(node, module_list):
    return node.module_ if node.module_ is not None else module_list[node.type]


def flatten_tree(tree, module_list=None, terrain_height=K.TERRAIN_HEIGHT):
    """Assemble one creature. ``tree`` is not modified."""
    nodes = tree.getNodes() if hasattr(tree, "getNodes") else list(tree)
    if module_list is None:
        module_list = tree.moduleList
    out = CreatureTable()
    slot_of = {}          # id(node) -> body slot, for nodes that produced a body
    handled = {}          # node index -> FIRST node with that index that went through create_component (built OR dropped)
    out.expressed = [-1] * len(nodes)

    def emit(pos_of_node, node, x, y, angle, mod):
        slot = out.n_bodies
        if isinstance(mod, Circular2D) or getattr(mod, "type", None) == "CIRCLE":
            out.shape.append(K.SHAPE_CIRCLE)
            out.hx.append(_f32(mod.radius))
            out.hy.append(0.0)
        else:
            out.shape.append(K.SHAPE_BOX)
            out.hx.append(_f32(mod.width / 2))
            out.hy.append(_f32(mod.height / 2))
        out.x0.append(_f32(x))
        out.y0.append(_f32(y))
        out.a0.append(_f32(angle))
        out.node_index.append(int(node.index))
        out.type_ref.append(int(node.type))
        c = node.controller
        out.ctrl.append((float(c.amplitude), float(c.phase), float(c.frequency), float(c.offset), float(c.i_state)))
        slot_of[id(node)] = slot
        out.expressed[pos_of_node] = slot
        return slot

    # pass 1: roots (create_robot, Modular2DEnv.py:521-526) — pose (5, TERRAIN_HEIGHT + 2), angle 0
    done = set()
    for k, node in enumerate(nodes):
        if node.parent == -1:
            mod = _module_of(node, module_list)
            if not mod.too_low(K.ROOT_Y, terrain_height):
                emit(k, node, K.ROOT_X, K.ROOT_Y, 0.0, mod)
            handled.setdefault(node.index, node)
            done.add(k)

    # pass 2: every other node, in list order, if its parent produced a body
    for k, node in enumerate(nodes):
        if k in done:
            continue
        parent = handled.get(node.parent)       # get_component_index: first handled node with that index
        if parent is None or id(parent) not in slot_of:
            continue                             # parent missing or dropped: node is never expressed
        ps = slot_of[id(parent)]
        pmod = parent.module_
        if pmod is None:
            raise IndexError("parent node without module_: the reference cannot place a connection site")
        px, py, pa = out.x0[ps], out.y0[ps], out.a0[ps]      # float32-rounded pose, as read back from Box2D
        con = node.parent_connection_coordinates
        if getattr(pmod, "type", None) == "CIRCLE":
            # circular_module.py:138-155 — unreachable with the stock encodings (circles are leaves)
            s_ang = con.value[0] * pmod.angle + pa
            site = (math.cos(s_ang + math.pi / 2) * pmod.radius + px,
                    math.sin(s_ang + math.pi / 2) * pmod.radius + py)
        else:
            site, s_ang = pmod.connection_site(con, px, py, pa)
        mod = _module_of(node, module_list)
        cx, cy = mod.child_placement(site, s_ang)
        handled.setdefault(node.index, node)
        if mod.too_low(cy, terrain_height):
            continue                             # dropped (simple_module.py:268-271, circular_module.py:186-189)
        slot = emit(k, node, cx, cy, 0 + s_ang, mod)
        # revolute joint parent -> child at the site (module_utility.py:7-33), child pose read back as float32
        bx, by, ba = out.x0[slot], out.y0[slot], out.a0[slot]
        dis_a = math.sqrt(math.pow(site[0] - px, 2) + math.pow(site[1] - py, 2))
        dis_b = math.sqrt(math.pow(site[0] - bx, 2) + math.pow(site[1] - by, 2))
        ang_a = s_ang - pa + math.pi / 2
        ang_b = ba - s_ang - math.pi / 2
        out.joint_parent.append(ps)
        out.anchor_a.append((_f32(math.cos(ang_a) * dis_a), _f32(math.sin(ang_a) * dis_a)))
        out.anchor_b.append((_f32(math.cos(ang_b) * dis_b), _f32(math.sin(ang_b) * dis_b)))
        out.lower.append(_JOINT_LOWER32)
        out.upper.append(_JOINT_UPPER32)
        out.max_torque.append(_f32(mod.torque))
    return out


@dataclass
class PopulationTable:
    """CSR-by-creature SoA table. ``body_off[c]..body_off[c+1]`` are creature c's
    bodies; its joints are ``body_off[c]-c .. body_off[c+1]-(c+1)`` (one per non-root body)."""
    body_off: np.ndarray
    shape: np.ndarray
    hx: np.ndarray
    hy: np.ndarray
    x0: np.ndarray
    y0: np.ndarray
    a0: np.ndarray
    node_index: np.ndarray
    type_ref: np.ndarray
    joint_parent: np.ndarray
    anchor_a: np.ndarray
    anchor_b: np.ndarray
    lower: np.ndarray
    upper: np.ndarray
    max_torque: np.ndarray
    ctrl: np.ndarray            # [n_bodies, 5] float64: amplitude, phase, frequency, offset, i_state

    @property
    def n_creatures(self):
        return len(self.body_off) - 1

    @property
    def n_bodies(self):
        return int(self.body_off[-1])

    def joint_off(self):
        return self.body_off - np.arange(len(self.body_off), dtype=self.body_off.dtype)

    def select(self, idx):
        """Sub-population (used for sharding by individual)."""
        idx = np.asarray(idx, dtype=np.int64)
        nb = (self.body_off[1:] - self.body_off[:-1])[idx]
        boff = np.zeros(len(idx) + 1, dtype=np.int32)
        np.cumsum(nb, out=boff[1:])

        def ranges(start, count):          # concatenation of arange(start[k], start[k] + count[k]) without a Python loop
            count = count.astype(np.int64)
            first = np.zeros(len(count) + 1, np.int64)
            np.cumsum(count, out=first[1:])
            return np.repeat(start.astype(np.int64) - first[:-1], count) + np.arange(first[-1], dtype=np.int64)

        bsel = ranges(self.body_off[:-1][idx], nb)
        jsel = ranges(self.joint_off()[:-1][idx], nb - 1)
        return PopulationTable(boff, self.shape[bsel], self.hx[bsel], self.hy[bsel], self.x0[bsel], self.y0[bsel],
                               self.a0[bsel], self.node_index[bsel], self.type_ref[bsel], self.joint_parent[jsel],
                               self.anchor_a[jsel], self.anchor_b[jsel], self.lower[jsel], self.upper[jsel],
                               self.max_torque[jsel], self.ctrl[bsel])


def pack(creatures):
    """list[CreatureTable] -> PopulationTable."""
    nb = np.array([c.n_bodies for c in creatures], dtype=np.int32)
    if (nb < 1).any():
        raise ValueError("every creature needs at least its root body")
    boff = np.zeros(len(creatures) + 1, dtype=np.int32)
    np.cumsum(nb, out=boff[1:])

    def cat(attr, dtype, width=None):
        vals = [v for c in creatures for v in getattr(c, attr)]
        if width is None:
            return np.asarray(vals, dtype=dtype).reshape(-1)
        return np.asarray(vals, dtype=dtype).reshape(-1, width)

    return PopulationTable(
        boff, cat("shape", np.uint8), cat("hx", np.float32), cat("hy", np.float32), cat("x0", np.float32),
        cat("y0", np.float32), cat("a0", np.float32), cat("node_index", np.int32), cat("type_ref", np.int16),
        cat("joint_parent", np.int16), cat("anchor_a", np.float32, 2), cat("anchor_b", np.float32, 2),
        cat("lower", np.float32), cat("upper", np.float32), cat("max_torque", np.float32),
        cat("ctrl", np.float64, 5))


def flatten_population(individuals, tree_depth=None):
    """Expand every genome (``genome.create``) and flatten it; accepts Individuals or Trees."""
    tables = []
    for ind in individuals:
        if hasattr(ind, "genome"):
            depth = tree_depth if tree_depth is not None else ind.tree_depth
            # the flattener only reads the tree: the L-system may hand out nodes that share the type's module / controller
            tree = ind.genome.create(depth, share=True) if isinstance(ind.genome, LSystem) else ind.genome.create(depth)
            tables.append(flatten_tree(tree, ind.genome.moduleList))
        else:
            tables.append(flatten_tree(ind))
    return pack(tables)
