"""Edge cases of the boundary: empty / ragged / maximum-size inputs and invalid tables.
CPU part runs on the oracle; the gpu-marked part checks the CUDA library against it."""
import math
import random

import numpy as np
import pytest

from gym_rem2d_b200 import Individual, constants as K, terrain
from gym_rem2d_b200.capi import Engine, Rem2dError
from gym_rem2d_b200.flatten import PopulationTable, flatten_population, flatten_tree, pack
from gym_rem2d_b200.modules import Circular2D, Connection, Standard2D
from gym_rem2d_b200.tree import Node, Tree
from oracle.oracle import OracleEngine


def tower(n_boxes, seed=0):
    """A chain of n boxes stacked through their 'top' sites, with alternating side branches of circles."""
    random.seed(seed)
    mods = [Standard2D() for _ in range(n_boxes)]
    for m in mods:
        m.width, m.height, m.angle = 0.5, 0.6, math.pi / 3
    t = Tree(mods)
    t.nodes = [Node(0, -1, 0, None, mods[0].controller, module_=mods[0])]
    for i in range(1, n_boxes):
        t.nodes.append(Node(i, i - 1, 0, Connection.top, mods[i].controller, module_=mods[i]))
    return t


def empty_table():
    z = lambda dt, *shape: np.zeros(shape, dt)
    return PopulationTable(np.zeros(1, np.int32), z(np.uint8, 0), z(np.float32, 0), z(np.float32, 0), z(np.float32, 0),
                           z(np.float32, 0), z(np.float32, 0), z(np.int32, 0), z(np.int16, 0), z(np.int16, 0),
                           z(np.float32, 0, 2), z(np.float32, 0, 2), z(np.float32, 0), z(np.float32, 0), z(np.float32, 0),
                           z(np.float64, 0, 5))


def _oracle():
    e = OracleEngine(threads=4)
    xs, ys = terrain.generate_terrain()
    e.set_terrain(ys, K.TERRAIN_STEP)
    return e


def test_maximum_size_creature_is_flattened_completely():
    c = flatten_tree(tower(44))
    assert c.n_bodies == 44 and len(c.joint_parent) == 43
    assert c.joint_parent == list(range(43))


def test_oracle_empty_and_ragged_population():
    e = _oracle()
    fit, ticks = e.evaluate(empty_table(), 100)
    assert len(fit) == 0 and len(ticks) == 0
    random.seed(1)
    single = Tree([Circular2D()])
    single.nodes = [Node(0, -1, 0, None, single.moduleList[0].controller, module_=single.moduleList[0])]
    pop = pack([flatten_tree(single), flatten_tree(tower(44)), flatten_tree(tower(2)), flatten_tree(single)])
    assert np.diff(pop.body_off).tolist() == [1, 44, 2, 1]
    fit, ticks = e.evaluate(pop, 300)
    assert ticks[0] == ticks[3] and fit[0] == fit[3]           # identical creatures, identical results
    assert np.all(ticks > 50) and np.all(np.isfinite(fit))


def test_oracle_rejects_invalid_tables():
    e = _oracle()
    pop = pack([flatten_tree(tower(3))])
    bad = pop.select([0])
    bad.joint_parent = bad.joint_parent.copy()
    bad.joint_parent[1] = 2                                       # parent must precede its child
    with pytest.raises(Rem2dError):
        e.upload(bad)
    with pytest.raises(Rem2dError):
        OracleEngine().upload(pop)                                # no terrain set


def test_env_api_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from gym_rem2d_b200 import env
    with pytest.raises(Rem2dError):
        env.BatchedModular2D()


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_gpu_empty_single_and_maximum_size_match_oracle():
    xs, ys = terrain.generate_terrain()
    g = Engine(device=0)
    g.set_terrain(ys, K.TERRAIN_STEP)
    o = _oracle()
    fit, ticks = g.evaluate(empty_table(), 100)
    assert len(fit) == 0
    random.seed(1)
    single = Tree([Circular2D()])
    single.nodes = [Node(0, -1, 0, None, single.moduleList[0].controller, module_=single.moduleList[0])]
    pop = pack([flatten_tree(single), flatten_tree(tower(44)), flatten_tree(tower(2)), flatten_tree(tower(33)), flatten_tree(single)])
    fg, tg = g.evaluate(pop, 400)
    fo, to = o.evaluate(pop, 400)
    assert np.array_equal(tg, to) and np.array_equal(fg, fo)
    # stepping path on the same ragged population, state compared tick by tick for a while
    g.upload(pop); o.upload(pop)
    for _ in range(40):
        g.step(1); o.step(1)
        sg, so = g.read_state(max_pairs=64), o.read_state(max_pairs=64)
        for k in ("pose", "vel", "joint_impulse", "touching_pairs", "limit_state", "n_contacts"):
            assert np.array_equal(sg[k], so[k]), k


@pytest.mark.gpu
def test_gpu_rejects_invalid_and_oversized_tables():
    xs, ys = terrain.generate_terrain()
    g = Engine(device=0)
    g.set_terrain(ys, K.TERRAIN_STEP)
    with pytest.raises(Rem2dError, match="bodies"):
        g.upload(pack([flatten_tree(tower(45))]))                 # larger than the largest capacity class
    bad = pack([flatten_tree(tower(3))])
    bad.joint_parent = bad.joint_parent.copy(); bad.joint_parent[1] = 2
    with pytest.raises(Rem2dError):
        g.upload(bad)
    with pytest.raises(Rem2dError):
        Engine(device=0).upload(pack([flatten_tree(tower(3))]))   # no terrain
    with pytest.raises(Rem2dError):
        Engine(device=99)


@pytest.mark.gpu
def test_gpu_env_api_matches_reference_signatures():
    from gym_rem2d_b200 import env
    random.seed(9)
    inds = [Individual.random(encoding="direct") for _ in range(8)]
    f1 = env.evaluate(inds[0])
    fp = env.evaluate_population(inds, as_torch=False)
    assert isinstance(f1, float) and f1 == fp[0]
    e = env.Modular2D()
    e.seed(4)
    tree = inds[1].genome.create(8)
    e.reset(tree=tree, module_list=inds[1].genome.moduleList)
    obs, reward, done, info = e.step(np.ones(4))
    assert obs == 0 and info == 0 and isinstance(reward, float) and done in (True, False)
    total = 0
    while not done and total < 400:
        obs, reward, done, info = e.step(None)
        total += 1
    assert done and reward == -100.0                              # wall of death reaches a random creature eventually
    e.close()
