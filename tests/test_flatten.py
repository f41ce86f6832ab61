"""P1/P2 pins: terrain and morphology tables vs fixtures recorded from the reference itself
(tests/golden/make_golden.py). Bit-exact: float32 tables compared as raw bits, float64 exactly."""
import os
import random

import numpy as np
import pytest

from gym_rem2d_b200 import Individual
from gym_rem2d_b200 import terrain
from gym_rem2d_b200.flatten import flatten_tree, pack, flatten_population


def test_terrain_rough_and_flat(golden_dir):
    g = np.load(os.path.join(golden_dir, "terrain.npz"))
    xs, ys = terrain.generate_terrain()
    assert np.array_equal(xs, g["rough_x"]) and np.array_equal(ys, g["rough_y"])
    fx, fy = terrain.flat_terrain()
    assert np.array_equal(fx, g["flat_x"]) and np.array_equal(fy, g["flat_y"])
    assert np.all(fy == 5.0)
    # the 199 edge fixtures are (x_i, y_i)-(x_{i+1}, y_{i+1}) in ascending x (Modular2DEnv.py:294-306)
    e = g["rough_edges"]
    assert e.shape == (199, 2, 2)
    assert np.array_equal(e[:, 0, 0], xs[:-1]) and np.array_equal(e[:, 1, 1], ys[1:])


def test_terrain_known_values():
    # SURVEY.md Appendix C
    xs, ys = terrain.generate_terrain()
    assert np.all(ys[:21] == 5.0)
    assert abs(ys[21] - 5.0434) < 1e-4 and abs(ys.min() - 4.2128) < 1e-4 and abs(ys.max() - 15.993) < 1e-3
    assert abs(xs[199] - 92.8667) < 1e-4


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


@pytest.mark.parametrize("enc", ["direct", "lsystem", "ce"])
def test_morphology_tables_bit_exact(enc, golden_dir):
    g = np.load(os.path.join(golden_dir, "morphology_%s.npz" % enc))
    n_mut = int(g["n_mutated"])
    mut = tuple(g["mut_args"])
    tables = []
    for i, seed in enumerate(g["seeds"]):
        random.seed(int(seed))
        ind = Individual.random(encoding=enc)
        tables.append(flatten_tree(ind.genome.create(ind.tree_depth), ind.genome.moduleList))
        if i < n_mut:
            for _ in range(3):
                ind.genome.mutate(*mut)
            tables.append(flatten_tree(ind.genome.create(ind.tree_depth), ind.genome.moduleList))
    pop = pack(tables)
    assert np.array_equal(pop.body_off, g["body_off"])
    for k in ("shape", "hx", "hy", "x0", "y0", "a0", "joint_parent", "anchor_a", "anchor_b", "lower", "upper",
              "max_torque", "node_index", "type_ref"):
        assert np.array_equal(_bits(getattr(pop, k)), _bits(g[k])), k
    assert np.array_equal(pop.ctrl, g["ctrl"])


def test_worked_example_appendix_b():
    """Un-mutated 0.2 x 0.8 root with the same module on 'top' (SURVEY.md Appendix B)."""
    from gym_rem2d_b200.modules import Standard2D, Connection
    from gym_rem2d_b200.tree import Tree, Node
    random.seed(0)
    m = Standard2D()
    t = Tree([m])
    t.nodes = [Node(0, -1, 0, None, m.controller, module_=m), Node(1, 0, 0, Connection.top, m.controller, module_=m)]
    c = flatten_tree(t)
    assert c.n_bodies == 2
    assert np.float32(c.y0[1]) == np.float32(7.8) and c.x0[1] == 5.0 and c.a0[1] == 0.0
    assert np.float32(c.anchor_a[0][1]) == np.float32(0.4)
    assert np.float32(c.anchor_b[0][1]) == np.float32(-0.40000019)
    # left/right children of an un-mutated box start exactly on the +-pi/2 limit
    t.nodes[1] = Node(1, 0, 0, Connection.left, m.controller, module_=m)
    c = flatten_tree(t)
    assert np.float32(c.a0[1]) == np.float32(-np.pi / 2)


def test_dropped_nodes_and_orphans():
    """A module whose centre is too low is dropped together with its whole sub-tree."""
    from gym_rem2d_b200.modules import Standard2D, Circular2D, Connection
    from gym_rem2d_b200.tree import Tree, Node
    random.seed(1)
    box = Standard2D()
    box.width, box.height, box.angle = 1.0, 1.0, np.pi     # left/right sites point straight down/up
    t = Tree([box])
    nodes = [Node(0, -1, 0, None, box.controller, module_=box)]
    # chain of boxes hanging downwards: 7 -> 6 -> 5 ... the second one violates y - sqrt(2) >= 5
    nodes.append(Node(1, 0, 0, Connection.left, box.controller, module_=box))
    nodes.append(Node(2, 1, 0, Connection.top, box.controller, module_=box))
    nodes.append(Node(3, 0, 0, Connection.top, Circular2D().controller, module_=Circular2D()))
    t.nodes = nodes
    c = flatten_tree(t)
    assert c.expressed[0] == 0 and c.expressed[3] >= 0
    assert c.expressed[2] == -1 or c.expressed[1] >= 0   # a child is never built without its parent
    assert len(c.joint_parent) == c.n_bodies - 1


def test_population_table_select_roundtrip():
    random.seed(5)
    inds = [Individual.random(encoding="lsystem") for _ in range(12)]
    pop = flatten_population(inds)
    assert pop.n_creatures == 12
    sub = pop.select([3, 7, 11])
    ref = flatten_population([inds[3], inds[7], inds[11]])
    for k in ("body_off", "shape", "hx", "x0", "a0", "joint_parent", "anchor_a", "ctrl"):
        assert np.array_equal(getattr(sub, k), getattr(ref, k)), k
    joff = pop.joint_off()
    assert joff[-1] == pop.n_bodies - pop.n_creatures


def test_lsystem_shared_expansion_gives_the_same_table():
    """flatten_population expands L-systems without the per-node deep copies (it only reads the tree): the table must equal
    the one built from the reference-style tree, for fresh and for mutated genomes."""
    import random
    from gym_rem2d_b200.ea import default_config
    from gym_rem2d_b200.individual import Individual
    from gym_rem2d_b200.modules import get_module_list
    from gym_rem2d_b200.flatten import flatten_population, flatten_tree, pack
    random.seed(11); np.random.seed(11)
    cfg = default_config(enc="lsystem")
    pop = [Individual.random(get_module_list(), cfg) for _ in range(120)]
    for rnd in range(3):
        fast = flatten_population(pop, 8)
        slow = pack([flatten_tree(ind.genome.create(8), ind.genome.moduleList) for ind in pop])
        for name in ("body_off", "shape", "hx", "hy", "x0", "y0", "a0", "node_index", "type_ref", "joint_parent", "anchor_a",
                     "anchor_b", "lower", "upper", "max_torque", "ctrl"):
            a, b = getattr(fast, name), getattr(slow, name)
            assert a.dtype == b.dtype and a.shape == b.shape and a.tobytes() == b.tobytes(), name
        for ind in pop:
            Individual.mutate(0.3, 0.3, 0.3, ind)
