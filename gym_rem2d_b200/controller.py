"""Per-module sine oscillator (reference: Controller/m_controller.py:4-58).

The draw order of ``random`` calls is the same as the reference's so a seeded run produces the
same parameters; the per-tick ``update`` is what the CUDA step kernel evaluates for every joint.
"""
import math
import random


class Controller:
    MAX_AMP = 1
    MAX_PHASE = 1
    MAX_OFFSET = math.pi
    MAX_FREQ = 0.1

    def __init__(self):
        self.i_state = 0
        self.output = 0
        # instance copies of the bounds, like the reference's __init__ (m_controller.py:9-12): part of the pickled state
        self.MAX_AMP = 1
        self.MAX_PHASE = 1
        self.MAX_OFFSET = math.pi
        self.MAX_FREQ = 0.1
        self.amplitude = random.uniform(0, self.MAX_AMP)
        self.phase = random.uniform(-self.MAX_PHASE, self.MAX_PHASE)
        self.frequency = random.uniform(-self.MAX_FREQ, self.MAX_FREQ)
        self.offset = random.uniform(-self.MAX_OFFSET, self.MAX_OFFSET)

    def update(self, input):
        """One tick: out = A sin(state + phase) + offset, state += f (m_controller.py:17-21)."""
        self.phase += input
        self.i_state += self.frequency
        self.output = self.amplitude * math.sin(self.i_state + self.phase) + self.offset
        return self.output

    def minMax(self, angle):
        """Clamp to the legal ranges; the offset range depends on the module's joint angle."""
        self.amplitude = min(max(self.amplitude, 0), self.MAX_AMP)
        self.phase = min(max(self.phase, -self.MAX_PHASE), self.MAX_PHASE)
        self.frequency = min(max(self.frequency, -self.MAX_FREQ), self.MAX_FREQ)
        # the reference tests the upper bound first and uses elif (m_controller.py:37-40)
        if self.offset > angle / 2:
            self.offset = angle / 2
        elif self.offset < -angle / 2:
            self.offset = -angle / 2

    def setControl(self, a, b, c, d, angle):
        """Set the four parameters from network outputs in [-1, 1] (m_controller.py:43-48)."""
        self.amplitude = ((a + 1.0) * 0.5) * self.MAX_AMP
        self.phase = b * self.MAX_PHASE
        self.offset = c * self.MAX_OFFSET
        self.frequency = d * self.MAX_FREQ
        self.minMax(angle)

    def mutate(self, mutationrate, sigma, angle):
        """Gaussian mutation; note the reference *adds* a sample centred on the value itself."""
        if random.uniform(0.0, 1.0) < mutationrate:
            self.amplitude += random.gauss(self.amplitude, sigma)
        if random.uniform(0.0, 1.0) < mutationrate:
            self.phase += random.gauss(self.phase, sigma)
        if random.uniform(0.0, 1.0) < mutationrate:
            self.frequency += random.gauss(self.frequency, sigma * 0.1)
        if random.uniform(0.0, 1.0) < mutationrate:
            self.offset += random.gauss(self.offset, sigma)
        self.minMax(angle)
