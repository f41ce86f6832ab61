"""Headless rasteriser of one creature: replaces the reference's pyglet viewer
(``Modular2D.render``, Modular2DEnv.py:655-738) with a numpy software renderer, so a creature of a batched run can be
looked at - and the physics debugged - on a machine without a display.

Same scene composition as the reference: 800x600 viewport at SCALE = 30 px per metre that scrolls with the root module
(Modular2DEnv.py:636-641, without the reference's 0.99 smoothing term), sky-coloured background, the terrain polygons
below the height field, the wall of death as a red vertical band, then the modules in creation order (boxes as filled
rotated rectangles with an outline, circles as discs with an outline; fill colour = viridis-like ramp of the module's
controller value when given, like COLOR_CONTROL). Input is the state read through the C-ABI (``rem2d_read_state``) plus
the flattened table (shapes / extents). ``write_png`` needs only zlib.
"""
import struct
import zlib

import numpy as np

from . import constants as K

SKY = (230, 230, 255)
GROUND = (102, 153, 76)
GROUND_DARK = (76, 127, 51)
WOD = (220, 40, 40)
OUTLINE = (30, 30, 30)


def _ramp(v):
    """Small viridis-like colour ramp for v in [0, 1]."""
    v = float(min(1.0, max(0.0, v)))
    stops = np.array([[68, 1, 84], [59, 82, 139], [33, 145, 140], [94, 201, 98], [253, 231, 37]], np.float64)
    x = v * (len(stops) - 1)
    i = min(int(x), len(stops) - 2)
    c = stops[i] + (stops[i + 1] - stops[i]) * (x - i)
    return tuple(int(round(q)) for q in c)


class Canvas:
    def __init__(self, width=K.VIEWPORT_W, height=K.VIEWPORT_H, scale=K.SCALE, scroll=(0.0, 0.0)):
        self.w, self.h, self.scale = width, height, float(scale)
        self.sx, self.sy = scroll
        self.img = np.empty((height, width, 3), np.uint8)
        self.img[:] = SKY
        ys, xs = np.mgrid[0:height, 0:width]
        # world coordinates of the pixel centres (y up)
        self.X = self.sx + (xs + 0.5) / self.scale
        self.Y = self.sy + (height - ys - 0.5) / self.scale

    def fill(self, mask, color):
        self.img[mask] = color

    def polygon(self, pts, color):
        """Filled convex polygon (world coordinates, counter-clockwise or clockwise)."""
        pts = np.asarray(pts, np.float64)
        lo, hi = pts.min(0), pts.max(0)
        if hi[0] < self.sx or lo[0] > self.sx + self.w / self.scale:
            return
        inside_pos = np.ones(self.X.shape, bool)
        inside_neg = np.ones(self.X.shape, bool)
        for i in range(len(pts)):
            x1, y1 = pts[i]
            x2, y2 = pts[(i + 1) % len(pts)]
            cr = (x2 - x1) * (self.Y - y1) - (y2 - y1) * (self.X - x1)
            inside_pos &= cr >= 0
            inside_neg &= cr <= 0
        self.fill(inside_pos | inside_neg, color)

    def disc(self, c, r, color):
        self.fill((self.X - c[0]) ** 2 + (self.Y - c[1]) ** 2 <= r * r, color)

    def ring(self, c, r, color, width_px=2):
        d = np.sqrt((self.X - c[0]) ** 2 + (self.Y - c[1]) ** 2)
        self.fill(np.abs(d - r) <= width_px / self.scale / 2 + 1e-9, color)

    def segment(self, a, b, color, width_px=2):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        ab = b - a
        t = np.clip(((self.X - a[0]) * ab[0] + (self.Y - a[1]) * ab[1]) / max(float(ab @ ab), 1e-18), 0.0, 1.0)
        d2 = (self.X - (a[0] + t * ab[0])) ** 2 + (self.Y - (a[1] + t * ab[1])) ** 2
        self.fill(d2 <= (width_px / self.scale / 2) ** 2, color)


def render_creature(table, creature, pose, terrain_y, wod=None, controller_values=None, width=K.VIEWPORT_W, height=K.VIEWPORT_H,
                    scroll=None):
    """RGB image (height x width x 3, uint8) of creature ``creature`` of ``table`` at the body poses ``pose``
    (the (n_bodies, 3) array of ``Engine.read_state()['pose']`` for the whole table)."""
    b0, b1 = int(table.body_off[creature]), int(table.body_off[creature + 1])
    p = np.asarray(pose, np.float64)[b0:b1]
    if scroll is None:      # Modular2DEnv.py:636-637
        scroll = (p[0, 0] - width / K.SCALE / 5, p[0, 1] - height / K.SCALE / 4)
    cv = Canvas(width, height, K.SCALE, scroll)
    ty = np.asarray(terrain_y, np.float64)
    for i in range(len(ty) - 1):            # terrain_poly: the quad under every edge (Modular2DEnv.py:303-309)
        x1, x2 = i * K.TERRAIN_STEP, (i + 1) * K.TERRAIN_STEP
        if x2 < cv.sx or x1 > cv.sx + width / K.SCALE:
            continue
        cv.polygon([(x1, ty[i]), (x2, ty[i + 1]), (x2, min(cv.sy, ty[i + 1]) - 1.0), (x1, min(cv.sy, ty[i]) - 1.0)],
                   GROUND if i % 2 == 0 else GROUND_DARK)
    if wod is not None:
        cv.fill(np.abs(cv.X - float(wod)) <= 1.5 / K.SCALE, WOD)
    for k in range(b1 - b0):
        x, y, a = p[k]
        col = _ramp(controller_values[k]) if controller_values is not None else (160, 160, 200)
        if table.shape[b0 + k] == K.SHAPE_CIRCLE:
            r = float(table.hx[b0 + k])
            cv.disc((x, y), r, col)
            cv.ring((x, y), r, OUTLINE)
            cv.segment((x, y), (x + r * np.cos(a), y + r * np.sin(a)), OUTLINE, 1)
        else:
            hx, hy = float(table.hx[b0 + k]), float(table.hy[b0 + k])
            c, s = np.cos(a), np.sin(a)
            corners = [(x + c * dx - s * dy, y + s * dx + c * dy) for dx, dy in ((-hx, -hy), (hx, -hy), (hx, hy), (-hx, hy))]
            cv.polygon(corners, col)
            for i in range(4):
                cv.segment(corners[i], corners[(i + 1) % 4], OUTLINE)
    return cv.img


def write_png(path, img):
    """Minimal PNG writer (8-bit RGB, zlib only)."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w, _ = img.shape
    raw = b"".join(b"\x00" + img[r].tobytes() for r in range(h))

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def dump_trajectory(engine, table, n_ticks, every=1):
    """State dump of a batched run for replay (8f N2): steps the uploaded population ``n_ticks`` ticks and returns
    {'pose': (frames, n_bodies, 3) float32, 'wod': (frames, n_creatures), 'alive': ..., 'tick': (frames,)}; frame 0 is the
    initial state. Any creature of the batch can then be rendered frame by frame with ``render_creature``."""
    frames = {"pose": [], "wod": [], "alive": [], "tick": []}

    def grab(t):
        st = engine.read_state()
        frames["pose"].append(st["pose"].copy()); frames["wod"].append(st["wod"].copy())
        frames["alive"].append(st["alive"].copy()); frames["tick"].append(t)
    grab(0)
    t = 0
    while t < n_ticks:
        step = min(every, n_ticks - t)
        engine.step(step)
        t += step
        grab(t)
    return {k: np.array(v) for k, v in frames.items()}
