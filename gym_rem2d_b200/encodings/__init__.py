"""Genotype -> phenotype encodings (stay in Python by design; they feed the flattener)."""
from .direct import DirectEncoding  # noqa: F401
from .lsystem import LSystem  # noqa: F401
from .network import NN_enc  # noqa: F401
