"""rem2d-b200: batched, B200-native evaluation of gym_rem2D modular creatures."""
from .individual import Individual, Encoding_Type, get_module_list  # noqa: F401
from .tree import Tree, Node  # noqa: F401
