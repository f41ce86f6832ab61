"""Headless rasteriser (8f N4): scene composition on synthetic poses (CPU), and a GPU smoke through Modular2D.render."""
import random
import zlib

import numpy as np
import pytest

from gym_rem2d_b200 import Individual, constants as K, render, terrain
from gym_rem2d_b200.flatten import flatten_population


def test_rasteriser_draws_terrain_modules_and_wall_of_death(tmp_path):
    random.seed(1)
    table = flatten_population([Individual.random(encoding="direct") for _ in range(3)])
    xs, ys = terrain.generate_terrain()
    pose = np.stack([table.x0, table.y0, table.a0], 1)
    img = render.render_creature(table, 1, pose, ys, wod=3.0)
    assert img.shape == (K.VIEWPORT_H, K.VIEWPORT_W, 3) and img.dtype == np.uint8
    cols = {tuple(c) for c in img.reshape(-1, 3)[::7]}
    assert render.SKY in cols and (render.GROUND in cols or render.GROUND_DARK in cols) and render.OUTLINE in cols
    # the root module is drawn where the viewport puts it: 1/5 from the left, 1/4 from the bottom
    px, py = int(K.VIEWPORT_W / 5), int(K.VIEWPORT_H - K.VIEWPORT_H / 4)
    assert tuple(img[py, px]) not in (render.SKY, render.GROUND, render.GROUND_DARK)
    # wall of death at x = 3: a red column left of the root (root x = 5 -> 2 m = 60 px to the left)
    assert tuple(img[50, px - 60]) == render.WOD
    path = tmp_path / "frame.png"
    render.write_png(path, img)
    data = open(path, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n" and b"IDAT" in data
    i = data.index(b"IDAT")
    n = int.from_bytes(data[i - 4:i], "big")
    raw = zlib.decompress(data[i + 4:i + 4 + n])
    assert len(raw) == K.VIEWPORT_H * (1 + 3 * K.VIEWPORT_W)


@pytest.mark.gpu
def test_single_creature_env_renders_and_dumps_a_trajectory(tmp_path):
    from gym_rem2d_b200.env import Modular2D
    random.seed(2)
    ind = Individual.random(encoding="lsystem")
    env = Modular2D()
    tree = ind.genome.create(ind.tree_depth)
    env.seed(4)
    env.reset(tree=tree, module_list=ind.genome.moduleList)
    first = env.render(mode="rgb_array")
    for _ in range(30):
        obs, reward, done, info = env.step(None)
    assert isinstance(reward, float) and not done and reward > 0
    later = env.render(mode="rgb_array")
    assert first.shape == later.shape and (first != later).any()
    out = env.render(mode="human", path=str(tmp_path / "f.png"))
    assert open(out, "rb").read()[:4] == b"\x89PNG"
    b = env._batched
    b.engine.reset()
    traj = render.dump_trajectory(b.engine, b.table, 20, every=5)
    assert traj["pose"].shape[0] == 5 and traj["tick"].tolist() == [0, 5, 10, 15, 20]
    assert (traj["pose"][0] != traj["pose"][-1]).any()
