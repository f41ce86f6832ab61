"""Scalar activation functions used by the cellular encoding and the CPPN stand-in.

Same function set and formulas as the reference's NeuralNetwork/activations.py:13-82 (which in turn
mirrors neat-python's), expressed as a name -> callable table.
"""
import math


def _scaled(fn, gain, lo=-60.0, hi=60.0):
    def f(z):
        return fn(max(lo, min(hi, gain * z)))
    return f


def _inv(z):
    try:
        return 1.0 / z
    except ArithmeticError:
        return 0.0


FUNCTIONS = {
    'sigmoid': _scaled(lambda z: 1.0 / (1.0 + math.exp(-z)), 5.0),
    'tanh': _scaled(math.tanh, 2.5),
    'sin': _scaled(math.sin, 5.0),
    'gauss': lambda z: math.exp(-5.0 * max(-3.4, min(3.4, z)) ** 2),
    'relu': lambda z: z if z > 0.0 else 0.0,
    'softplus': _scaled(lambda z: 0.2 * math.log(1 + math.exp(z)), 5.0),
    'identity': lambda z: z,
    'clamped': lambda z: max(-1.0, min(1.0, z)),
    'inv': _inv,
    'log': lambda z: math.log(max(1e-7, z)),
    'exp': _scaled(math.exp, 1.0),
    'abs': abs,
    'hat': lambda z: max(0.0, 1 - abs(z)),
    'square': lambda z: z ** 2,
    'cube': lambda z: z ** 3,
}

# order used by random.choice in the cellular encoding (Cellular_Encoding.py:35-50)
CE_ORDER = ['sigmoid', 'tanh', 'sin', 'gauss', 'relu', 'softplus', 'identity', 'clamped', 'inv',
            'log', 'exp', 'abs', 'hat', 'square', 'cube']

# order of ``activation_options`` in NeuralNetwork/config:25
CPPN_ORDER = ['abs', 'clamped', 'cube', 'exp', 'gauss', 'hat', 'identity', 'inv', 'log', 'relu',
              'sigmoid', 'sin', 'softplus', 'square', 'tanh']


class Activation:
    """Picklable handle on an activation, one instance per name (the reference stores the bare functions of
    NeuralNetwork/activations.py, which pickle by name; refpickle.dump writes these handles under those names)."""
    __slots__ = ("name",)
    _instances = {}
    reference_names = False        # set by refpickle.reference_class_paths()

    def __new__(cls, name):
        inst = cls._instances.get(name)
        if inst is None:
            if name not in FUNCTIONS:
                raise TypeError("No such activation function: {0!r}".format(name))
            inst = object.__new__(cls)
            inst.name = name
            cls._instances[name] = inst
        return inst

    def __call__(self, z):
        return FUNCTIONS[self.name](z)

    def __reduce__(self):
        if Activation.reference_names:
            return self.name + "_activation"          # global NeuralNetwork.activations.<name>_activation
        return (Activation, (self.name,))

    def __deepcopy__(self, memo):
        return self

    def __copy__(self):
        return self
