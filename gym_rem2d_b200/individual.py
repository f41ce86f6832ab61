"""Individual = genome holder + factory by encoding name (reference: REM2D_main.py:79-138)."""
from enum import Enum

from .encodings.direct import DirectEncoding
from .encodings.lsystem import LSystem
from .encodings.network import NN_enc
from .modules import get_module_list  # noqa: F401  (re-exported like REM2D_main.get_module_list)


class Encoding_Type(Enum):
    DIRECT = 0
    LSYSTEM = 1
    NEURAL_NETWORK = 2
    CELLULAR_ENCODING = 3


_FACTORIES = {
    'direct': (Encoding_Type.DIRECT, lambda ml, cfg: DirectEncoding(ml, cfg) if cfg is not None else DirectEncoding(ml)),
    'lsystem': (Encoding_Type.LSYSTEM, lambda ml, cfg: LSystem(ml, cfg) if cfg is not None else LSystem(ml)),
    'cppn': (Encoding_Type.NEURAL_NETWORK, lambda ml, cfg: NN_enc(ml, "CPPN", config=cfg)),
    'ce': (Encoding_Type.CELLULAR_ENCODING, lambda ml, cfg: NN_enc(ml, "CE", config=cfg)),
}


class Individual:
    def __init__(self):
        self.genome = None
        self.fitness = 0

    @staticmethod
    def random(moduleList=None, config=None, encoding='lsystem'):
        """Random individual of the given encoding ('direct' | 'lsystem' | 'cppn' | 'ce').

        With a config the encoding name and tree depth come from it (REM2D_main.py:96-115), else
        ``encoding`` is used and the depth is 8 (REM2D_main.py:116-133).
        """
        self = Individual()
        if moduleList is None:
            moduleList = get_module_list()
        name = config['encoding']['type'] if config is not None else encoding
        if name not in _FACTORIES:
            raise Exception("Could not find specified encoding type, please use 'direct','lsystem','cppn' or 'ce'")
        self.ENCODING_TYPE, make = _FACTORIES[name]
        self.genome = make(moduleList, config)
        self.tree_depth = int(config['morphology']['max_depth']) if config is not None else 8
        self.genome.create(self.tree_depth)
        self.fitness = 0
        return self

    def mutate(MORPH_MUTATION_RATE, MUTATION_RATE, MUT_SIGMA, self):
        # argument order as registered with the DEAP toolbox (REM2D_main.py:135-136,250)
        self.genome.mutate(MORPH_MUTATION_RATE, MUTATION_RATE, MUT_SIGMA)
