"""Phenotype blueprint handed from the encodings to the evaluation path.

Same data model as the reference's Tree.py:5-26 (field names kept so pickles and tooling that
poke at ``node.parent_connection_coordinates`` etc. keep working).
"""


class Tree:
    def __init__(self, moduleList, controller=None):
        self.nodes = []
        self.moduleList = moduleList

    def getNodes(self):
        return self.nodes


class Node:
    def __init__(self, index, parent, type, parent_connection_coordinates, controller=None,
                 component=None, module_=None):
        self.index = index
        self.type = type
        self.parent = parent
        self.parent_connection_coordinates = parent_connection_coordinates
        self.controller = controller      # decentralised sine controller of this module
        self.expressed = False            # set by the flattener, like create_robot does
        self.component = component        # body slot in the flattened table (or None if dropped)
        self.module_ = module_

    def __bool__(self):
        return self.expressed
