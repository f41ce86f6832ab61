"""Module types of the modular creatures: parameter holders + attachment geometry.

Reference: gym_rem2D/morph/simple_module.py (Standard2D), circular_module.py (Circular2D),
abstract_module.py:38-52 (``available``). The reference classes also create pybox2d bodies; here the
geometry is exposed as plain functions of numbers (``connection_site``, ``child_placement``) that the
flattener (flatten.py) turns into the SoA table consumed by the CUDA library. No physics objects.
"""
import math
import random
from enum import Enum

import numpy as np

from .controller import Controller


class Connection(Enum):
    """Connection sites of the rectangular module (simple_module.py:21-25)."""
    left = (-1., 0., 0.)
    right = (1., 0., 0.)
    top = (0., 1.0, 0.)


class CircularConnection(Enum):
    """Declared by the reference (circular_module.py:23-27) but never reachable: Circular2D does not
    set ``connection_type`` so it has no connection sites and is always a leaf."""
    left = (1., 0., 0.)
    right = (-1., 0., 0.)
    top = (0., 1.0, 0.)


class Module:
    connection_type = None
    _children = None

    @property
    def available(self):
        """Free connection sites, in enum declaration order (abstract_module.py:38-52)."""
        if not self.connection_type:
            return []
        return [c for c in self.connection_type if c not in self._children]


class Standard2D(Module):
    """Rectangle module (simple_module.py:27-92)."""
    type = "SIMPLE"
    MAX_HEIGHT = 1.0
    MIN_HEIGHT = 0.5
    MAX_WIDTH = 1.0
    MIN_WIDTH = 0.5
    MAX_ANGLE = math.pi
    MIN_ANGLE = 0

    def __init__(self, theta=0, size=(0.1, 0.1, 0.0)):
        # the instance carries every attribute the reference's __init__ sets (simple_module.py:29-53), bounds included, so
        # that a pickled module is a complete state for the reference's class as well (refpickle.py)
        self.theta = theta % 2
        self.size = np.array(size)
        self.position = np.array([0., self.size[2] / 2. + 0.002, 0.])
        self.connection_type = Connection
        self._children = {}
        self.controller = Controller()
        self.width = 0.2
        self.height = 0.8
        self.angle = math.pi / 2
        self.type = "SIMPLE"
        self.MAX_HEIGHT = 1.0
        self.MIN_HEIGHT = 0.5
        self.MAX_WIDTH = 1.0
        self.MIN_WIDTH = 0.5
        self.MAX_ANGLE = math.pi
        self.MIN_ANGLE = 0
        self.torque = 50

    def limitWH(self):
        self.height = _clamp_like_ref(self.height, self.MIN_HEIGHT, self.MAX_HEIGHT)
        self.width = _clamp_like_ref(self.width, self.MIN_WIDTH, self.MAX_WIDTH)
        self.angle = _clamp_like_ref(self.angle, self.MIN_ANGLE, self.MAX_ANGLE)

    def mutate(self, MORPH_MUTATION_RATE, MUTATION_RATE, MUT_SIGMA):
        if random.uniform(0, 1) < MORPH_MUTATION_RATE:
            self.width = random.gauss(self.width, MUT_SIGMA)
        if random.uniform(0, 1) < MORPH_MUTATION_RATE:
            self.height = random.gauss(self.height, MUT_SIGMA)
        if random.uniform(0, 1) < MORPH_MUTATION_RATE:
            self.angle = random.gauss(self.angle, MUT_SIGMA * math.pi)
        self.limitWH()
        if self.controller:
            self.controller.mutate(MUTATION_RATE, MUT_SIGMA, self.angle)

    def setMorph(self, val1, val2, val3):
        # the reference derives BOTH width and height from val1 (simple_module.py:87-92)
        self.width = (val1 * 0.5 * (self.MAX_WIDTH - self.MIN_WIDTH)) + 0.5 * (self.MAX_WIDTH - self.MIN_WIDTH)
        self.height = (val1 * 0.5 * (self.MAX_HEIGHT - self.MIN_HEIGHT)) + 0.5 * (self.MAX_HEIGHT - self.MIN_HEIGHT)
        self.angle = self.MIN_ANGLE + (((val3 + 1.0) * 0.5) * (self.MAX_ANGLE - self.MIN_ANGLE))
        self.limitWH()

    # ---- geometry (doubles in, doubles out; the flattener does the float32 round trips) ----
    def connection_site(self, con, parent_x, parent_y, parent_angle):
        """World position + orientation of site ``con`` on a parent rectangle whose Box2D pose is
        (parent_x, parent_y, parent_angle). Restates simple_module.py:147-199."""
        if con is None:
            con = Connection.left
        theta = con.value[0] * self.angle + math.pi / 2
        while theta > 2 * math.pi:      # unreachable for angle <= pi (the reference would NameError)
            theta -= 2 * math.pi
        sx = -1. if (0.5 * math.pi < theta < 1.5 * math.pi) else 1.
        sy = -1. if (math.pi < theta < 2 * math.pi) else 1.
        # ray from the centre in direction theta against the top/bottom edge ...
        if 2 * math.sin(theta) == 0:
            p1 = (10000, 10000)
        else:
            p1 = ((self.height * math.cos(theta)) / (2 * math.sin(theta)) * sy, self.height / 2 * sy)
        # ... and against the left/right edge
        if 2 * math.cos(theta) == 0:
            p2 = (10000, 10000)
        else:
            p2 = (self.width / 2 * sx, (self.width * math.sin(theta)) / (2 * math.cos(theta)) * sx)
        d1 = math.sqrt(math.pow(p1[0], 2) + math.pow(p1[1], 2))
        d2 = math.sqrt(math.pow(p2[0], 2) + math.pow(p2[1], 2))
        dist = d2 if d2 < d1 else d1
        gx = (math.cos(parent_angle + theta) * dist) + parent_x
        gy = (math.sin(parent_angle + theta) * dist) + parent_y
        return (gx, gy), parent_angle + theta - math.pi / 2

    def child_placement(self, site_pos, site_angle):
        """Centre of a rectangle attached at a site, and the too-low test (simple_module.py:255-271)."""
        x = math.cos(site_angle + 0 + math.pi / 2) * self.height / 2 + site_pos[0]
        y = math.sin(site_angle + 0 + math.pi / 2) * self.height / 2 + site_pos[1]
        return x, y

    def too_low(self, y, terrain_height):
        return y - math.sqrt(math.pow(self.width, 2) + math.pow(self.height, 2)) < terrain_height


class Circular2D(Module):
    """Circle module (circular_module.py:29-85). Never has children."""
    type = "CIRCLE"
    MIN_RADIUS = 0.25
    MAX_RADIUS = 0.5
    MIN_ANGLE = math.pi / 4
    MAX_ANGLE = math.pi * 2

    def __init__(self, theta=0, size=(0.1, 0.1, 0.0)):
        # every attribute of the reference's __init__ (circular_module.py:31-53) except connection_axis / orientation, which
        # are gym_rem (3-D tree) objects only touched by the dead 3-D methods of the reference class
        self.theta = theta % 2
        self.size = np.array(size)
        self.position = np.array([0., self.size[2] / 2. + 0.002, 0.])
        self._children = {}
        self.controller = Controller()
        self.radius = 0.25
        self.angle = math.pi / 2
        self.type = "CIRCLE"
        self.MIN_RADIUS = 0.25
        self.MAX_RADIUS = 0.5
        self.MIN_ANGLE = math.pi / 4
        self.MAX_ANGLE = math.pi * 2
        self.torque = 50

    def limitWH(self):
        self.radius = _clamp_like_ref(self.radius, self.MIN_RADIUS, self.MAX_RADIUS)
        self.angle = _clamp_like_ref(self.angle, self.MIN_ANGLE, self.MAX_ANGLE)

    def mutate(self, MORPH_MUTATION_RATE, MUTATION_RATE, MUT_SIGMA):
        if random.uniform(0, 1) < MORPH_MUTATION_RATE:
            self.radius = random.gauss(self.radius, MUT_SIGMA)
        if random.uniform(0, 1) < MORPH_MUTATION_RATE:
            self.angle = random.gauss(self.angle, MUT_SIGMA * math.pi)
        self.limitWH()
        if self.controller is not None:
            self.controller.mutate(MUTATION_RATE, MUT_SIGMA, self.angle)

    def setMorph(self, val1, val2, val3):
        self.radius = val1 + 1.5
        self.angle = self.MIN_ANGLE + (((val3 + 1.0) * 0.5) * (self.MAX_ANGLE - self.MIN_ANGLE))
        self.limitWH()

    def child_placement(self, site_pos, site_angle):
        """Centre of a circle attached at a site (circular_module.py:172-176)."""
        x = math.cos(site_angle + math.pi / 2) * self.radius + site_pos[0]
        y = math.sin(site_angle + math.pi / 2) * self.radius + site_pos[1]
        return x, y

    def too_low(self, y, terrain_height):
        return y - self.radius < terrain_height


def _clamp_like_ref(v, lo, hi):
    """``if v > hi: hi elif v < lo: lo`` — keeps the reference's value *types* (e.g. int 0)."""
    if v > hi:
        return hi
    elif v < lo:
        return lo
    return v


def get_module_list():
    """4 rectangles then 4 circles, each with a random controller (REM2D_main.py:69-77)."""
    return [Standard2D() for _ in range(4)] + [Circular2D() for _ in range(4)]
