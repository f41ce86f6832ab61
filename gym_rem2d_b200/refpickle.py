"""On-disk compatibility with the reference's pickles.

``run2D.run_deap`` checkpoints with plain ``pickle.dump`` of its own objects (REM2D_main.py:311-329): ``s_pop<i>`` = list of
``REM2D_main.Individual``, ``s_`` = ``DataAnalysis.FitnessData`` (DataAnalysis.py:39-56), ``s_elite<i>`` = one Individual;
``Experiments/Load_Best.py:7-37`` and ``DataAnalysis.load_data`` read them back with ``pickle.load``. A pickle names every
class by module path, so files written by this package would name ``gym_rem2d_b200.*`` classes and be unreadable there.

``dump`` writes the SAME object graph under the reference's class paths (the classes here keep the reference's attribute
names, so the reference's own classes accept the state); ``load`` reads either flavour into this package's classes.
The CPPN genome is the documented exception: the reference's is a neat-python object (``NeuralNetwork.NEAT_NN``), ours a
stand-in — CPPN individuals are written under this package's paths and a reference process cannot load them.
"""
import contextlib
import io
import pickle
import sys
import types

from . import controller, ea, individual, modules, tree
from .encodings import activations, cellular, direct, lsystem, network

# (our class, reference module, reference class name)
CLASS_MAP = [
    (individual.Individual, "REM2D_main", "Individual"),
    (individual.Encoding_Type, "REM2D_main", "Encoding_Type"),
    (ea.FitnessData, "DataAnalysis", "FitnessData"),
    (tree.Tree, "Tree", "Tree"),
    (tree.Node, "Tree", "Node"),
    (controller.Controller, "Controller.m_controller", "Controller"),
    (modules.Module, "gym_rem2D.morph.abstract_module", "Module"),
    (modules.Standard2D, "gym_rem2D.morph.simple_module", "Standard2D"),
    (modules.Connection, "gym_rem2D.morph.simple_module", "Connection"),
    (modules.Circular2D, "gym_rem2D.morph.circular_module", "Circular2D"),
    (modules.CircularConnection, "gym_rem2D.morph.circular_module", "Connection"),
    (direct.DirectNode, "Encodings.direct_encoding", "DirectNode"),
    (direct.DirectTree, "Encodings.direct_encoding", "DirectTree"),
    (direct.DirectEncoding, "Encodings.direct_encoding", "DirectEncoding"),
    (lsystem.C_Module, "Encodings.lsystem", "C_Module"),
    (lsystem.Rule, "Encodings.lsystem", "Rule"),
    (lsystem.LSystem, "Encodings.lsystem", "LSystem"),
    (cellular.Scheme, "Encodings.cellular_encoding", "Scheme"),
    (cellular.Link, "Encodings.cellular_encoding", "Link"),
    (cellular.Cell, "Encodings.cellular_encoding", "Cell"),
    (cellular.CE, "Encodings.cellular_encoding", "CE"),
    (network.C_Module, "Encodings.network_encoding", "C_Module"),
    (network.NN_enc, "Encodings.network_encoding", "NN_enc"),
    (network.NETWORK_TYPE, "Encodings.network_encoding", "NETWORK_TYPE"),
]
_REF_TO_OURS = {(m, n): c for c, m, n in CLASS_MAP}


@contextlib.contextmanager
def reference_class_paths():
    """Inside the block the classes of CLASS_MAP answer to the reference's module paths: pickle names a class by
    ``__module__`` / ``__qualname__`` and verifies that ``sys.modules[module].name`` is that class, so both are switched
    (alias modules are registered only if the real reference module is not loaded) and restored afterwards."""
    saved_attrs, saved_modules = [], {}
    try:
        for cls, mod, name in CLASS_MAP:
            saved_attrs.append((cls, cls.__module__, cls.__qualname__, cls.__name__))
            parts = mod.split(".")
            for i in range(1, len(parts) + 1):
                key = ".".join(parts[:i])
                if key not in saved_modules:
                    saved_modules[key] = sys.modules.get(key)
                    m = types.ModuleType(key)
                    m.__path__ = []
                    sys.modules[key] = m
            setattr(sys.modules[mod], name, cls)
            cls.__module__, cls.__qualname__, cls.__name__ = mod, name, name
        # activation handles are written as the reference's module-level functions NeuralNetwork.activations.<name>_activation
        act_mod = "NeuralNetwork.activations"
        for key in ("NeuralNetwork", act_mod):
            saved_modules[key] = sys.modules.get(key)
            m = types.ModuleType(key)
            m.__path__ = []
            sys.modules[key] = m
        for name in activations.FUNCTIONS:
            setattr(sys.modules[act_mod], name + "_activation", activations.Activation(name))
        saved_attrs.append((activations.Activation, activations.Activation.__module__, activations.Activation.__qualname__,
                            activations.Activation.__name__))
        activations.Activation.__module__ = act_mod
        activations.Activation.reference_names = True
        yield
    finally:
        activations.Activation.reference_names = False
        for cls, mod, qual, name in saved_attrs:
            cls.__module__, cls.__qualname__, cls.__name__ = mod, qual, name
        for key, old in saved_modules.items():
            if old is None:
                sys.modules.pop(key, None)
            else:
                sys.modules[key] = old


def dumps(obj, protocol=2):
    """Pickle ``obj`` (population list, Individual, FitnessData) under the reference's class paths. Protocol 2 is what every
    Python 3 the reference ran on can read."""
    buf = io.BytesIO()
    with reference_class_paths():
        pickle.Pickler(buf, protocol=protocol).dump(obj)
    return buf.getvalue()


def dump(obj, path, protocol=2):
    data = dumps(obj, protocol)
    with open(path, "wb") as f:
        f.write(data)


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == "NeuralNetwork.activations" and name.endswith("_activation"):
            return activations.Activation(name[:-len("_activation")])
        cls = _REF_TO_OURS.get((module, name))
        return cls if cls is not None else super().find_class(module, name)


def loads(data):
    """Unpickle a checkpoint written by the reference OR by this package into this package's classes."""
    return _Unpickler(io.BytesIO(data)).load()


def load(path):
    with open(path, "rb") as f:
        return loads(f.read())
