"""Terrain height-field of the REM2D environment.

Restates ``Modular2D.seed`` + ``_generate_terrain(hardcore=False)`` (Modular2DEnv.py:171-173,188-310):
200 vertices at x_i = i * 14/30; a smoothed random walk after a 20-vertex flat start pad. Because
``evaluate`` reseeds with 4 before every reset (REM2D_main.py:358-359) every creature of every
generation sees the same 199 edges, so the table is built once on the host and shared.

``np_random`` restates gym 0.18's ``gym.utils.seeding.np_random`` (third-party, not in the reference
tree): RandomState seeded with the little-endian uint32 words of sha512(str(seed))[:8].
"""
import hashlib
import struct

import numpy as np

from . import constants as K


def np_random(seed):
    digest = hashlib.sha512(str(seed).encode('utf8')).digest()[:8]
    words = struct.unpack('<2I', digest)
    big = words[0] + (words[1] << 32)
    ints = []
    while big > 0:
        big, lo = divmod(big, 2 ** 32)
        ints.append(lo)
    rng = np.random.RandomState()
    rng.seed(ints or [0])
    return rng


def generate_terrain(seed=K.TERRAIN_SEED, max_perturbance=K.MAX_PERTURBANCE_TERRAIN, rng=None):
    """Returns (terrain_x[200], terrain_y[200]) as float64 arrays, and leaves ``rng`` advanced exactly
    as the reference leaves ``env.np_random`` after ``_generate_terrain`` (clouds not included)."""
    if rng is None:
        rng = np_random(seed)
    n = K.TERRAIN_LENGTH
    xs = np.empty(n)
    ys = np.empty(n)
    velocity = 0.0
    y = K.TERRAIN_HEIGHT
    counter = K.TERRAIN_STARTPAD
    hold = False                      # 'oneshot': the vertex after a segment redraw repeats y
    for i in range(n):
        xs[i] = i * K.TERRAIN_STEP
        if not hold:
            velocity = 0.5 * velocity + 0.01 * np.sign(K.TERRAIN_HEIGHT - y)
            if i > K.TERRAIN_STARTPAD:
                amp = max_perturbance / K.TERRAIN_LENGTH * i
                velocity += rng.uniform(-amp, amp) / K.SCALE
            y += velocity
        hold = False
        ys[i] = y
        counter -= 1
        if counter == 0:
            counter = rng.randint(K.TERRAIN_GRASS / 2, K.TERRAIN_GRASS)
            hold = True
    return xs, ys


def flat_terrain():
    """``MAX_PERTURBANCE_TERRAIN = 0`` variant: y == 5.0 everywhere."""
    xs, ys = generate_terrain(max_perturbance=0)
    return xs, ys
