"""Physics / environment constants of the REM2D evaluation path.

Mirrors the module-level globals of the reference env (Modular2DEnv.py:26-66) and the fixture
constants of the two module types (simple_module.py:286-293, circular_module.py:191-198).
They are *not* configurable in the reference either.
"""
import math

FPS = 50                       # Modular2DEnv.py:26
SCALE = 30.0                   # Modular2DEnv.py:27
VELOCITY_ITERATIONS = 6 * 30   # Modular2DEnv.py:634
POSITION_ITERATIONS = 2 * 30   # Modular2DEnv.py:634
MAX_PERTURBANCE_TERRAIN = 24   # Modular2DEnv.py:34 (0 => flat terrain)
WOD_SPEED = 0.04               # Modular2DEnv.py:52 (per tick)
VIEWPORT_W = 800               # Modular2DEnv.py:54
VIEWPORT_H = 600               # Modular2DEnv.py:55
TERRAIN_STEP = 14 / SCALE      # Modular2DEnv.py:57
TERRAIN_LENGTH = 200           # Modular2DEnv.py:58
TERRAIN_HEIGHT = VIEWPORT_H / SCALE / 4   # Modular2DEnv.py:59  (= 5.0)
TERRAIN_GRASS = 10             # Modular2DEnv.py:60
TERRAIN_STARTPAD = 20          # Modular2DEnv.py:61
TERRAIN_FRICTION = 2.5         # Modular2DEnv.py:62
MODULE_FRICTION = 0.1          # simple_module.py:289
MODULE_DENSITY = 1.0           # simple_module.py:288
MODULE_RESTITUTION = 0.0       # simple_module.py:290
MODULE_TORQUE = 50.0           # simple_module.py:52
P_GAIN = 1.9                   # Modular2DEnv.py:601
JOINT_LOWER = -math.pi / 2     # module_utility.py:28
JOINT_UPPER = math.pi / 2      # module_utility.py:29
ROOT_X = 5.0                   # Modular2DEnv.py:430
ROOT_Y = TERRAIN_HEIGHT + 2    # Modular2DEnv.py:431
GRAVITY_Y = -10.0              # pybox2d b2World() default gravity
EVALUATION_STEPS = 10000       # REM2D_main.py:350
ENV_LENGTH = 100               # REM2D_main.py:350
TERRAIN_SEED = 4               # REM2D_main.py:358
TIME_LIMIT_STEPS = 240 * 20    # gym_rem2D/__init__.py:7 (gym TimeLimit wrapper)

SHAPE_BOX = 0
SHAPE_CIRCLE = 1
