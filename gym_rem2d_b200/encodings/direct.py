"""Direct tree encoding (reference: Encodings/Direct_Encoding.py:7-140).

The genome *is* the module tree. Mutation walks the tree, may drop sub-trees, may attach new random
modules to free connection sites, and perturbs module/controller parameters. The sequence of
``random`` draws matches the reference so a seeded run yields the same trees (pinned by
tests/golden/morphology_direct.json, generated from the reference itself).
"""
import copy
import random

from .. import tree as _tree
from ..controller import Controller


class DirectNode(_tree.Node):
    def __init__(self, index, parent, type, orientation, control, module_):
        super().__init__(index, parent, type, orientation, control, module_=copy.deepcopy(module_))
        self.availableConnections = self.module_.available
        self.children = []

    def addChild(self, module, index, parent, moduleRef, moduleController, parentConnectionSite):
        index += 1
        self.children.append(DirectNode(index, parent, moduleRef, parentConnectionSite, moduleController, module))
        self.availableConnections.remove(parentConnectionSite)
        return index


class DirectTree(_tree.Tree):
    def __init__(self, module_list):
        super().__init__(module_list)
        control = Controller()
        self.index = 0
        self.tree_nodes = [DirectNode(self.index, -1, 0, None, control, copy.deepcopy(module_list[0]))]

    def getNodes(self):
        """Pre-order flattening (parents before children), cached in ``self.nodes``."""
        out = []
        stack = [self.tree_nodes[0]]
        while stack:
            n = stack.pop()
            out.append(n)
            stack.extend(reversed(n.children))
        self.nodes = out
        return self.nodes


class DirectEncoding:
    def __init__(self, moduleList, config=None):
        self.moduleList = moduleList
        self.tree = DirectTree(moduleList)
        self.n_modules = 1
        if config is not None:
            self.maxDepth = int(config['morphology']['max_depth'])
            self.maxModules = int(config['morphology']['max_size'])
        else:
            self.maxDepth = 8
            self.maxModules = 20
        for _ in range(5):
            self.mutate(0.5, 0.5, 0.5)

    def create(self, treedepth):
        """The tree is the genome; only the controller phases are rewound (Direct_Encoding.py:61-71)."""
        for node in self.tree.nodes:
            node.controller.i_state = 0
        return self.tree

    def countModules(self):
        n = 0
        stack = [self.tree.tree_nodes[0]]
        while stack:
            node = stack.pop()
            n += 1
            stack.extend(node.children)
        self.n_modules = n

    def mutateNode(self, node, morphMutationRate, mutationRate, sigma, depth):
        self.countModules()
        # NB: like the reference this iterates the live child list while removing from it, so the
        # sibling after a removed child is skipped in this pass (Direct_Encoding.py:86-100).
        for mod in node.children:
            if random.uniform(0, 1) < float(morphMutationRate) / float(2) / float(self.n_modules):
                if depth != 0:
                    node.availableConnections.append(mod.parent_connection_coordinates)
                    node.children.remove(mod)
                    self.countModules()
            else:
                self.mutateNode(mod, morphMutationRate, mutationRate, sigma, depth + 1)
        # same live-iteration quirk: addChild removes ``con`` from the list being iterated
        for con in node.availableConnections:
            self.countModules()
            if (self.n_modules < self.maxModules and depth < self.maxDepth
                    and random.uniform(0, 1) < morphMutationRate / float(self.n_modules)):
                type = random.randint(0, len(self.moduleList) - 1)
                newModule = self.moduleList[type]
                moduleController = Controller()
                self.tree.index = node.addChild(copy.deepcopy(newModule), self.tree.index, node.index,
                                                type, moduleController, con)
        node.module_.mutate(morphMutationRate, mutationRate, sigma)
        node.controller.mutate(mutationRate, sigma, node.module_.angle)

    def reassignIndices(self):
        counter = [0]

        def walk(node):
            node.index = counter[0]
            counter[0] += 1
            for ch in node.children:
                walk(ch)
                ch.parent = node.index
        walk(self.tree.tree_nodes[0])
        self.index = counter[0]

    def mutate(self, morphMutationRate, mutationRate, sigma):
        self.mutateNode(self.tree.tree_nodes[0], morphMutationRate, mutationRate, sigma, 0)
        self.countModules()
        self.reassignIndices()
