"""Parametric L-system encoding (reference: Encodings/LSystem.py:19-199).

One rewriting rule per module type: ``A := A[children]``. ``create`` rewrites the axiom (rule 0)
``treeDepth`` times in parallel and converts the symbol tree to a Tree of Nodes in pre-order. All
nodes of one type share the type's module + controller parameters (deep copies).
"""
import copy
import random

from .. import tree as _tree


class C_Module:
    """Symbol of the L-system string (LSystem.py:19-30)."""

    def __init__(self, index, module, moduleRef):
        self.index = index
        self.parent = -1
        self.moduleRef = moduleRef
        self.availableConnections = copy.deepcopy(module.available)
        self.children = []
        self.theta = -1
        self.parentConnectionSite = None
        self.handled = False


def _random_product(rule_module, moduleList):
    con = random.choice(rule_module.availableConnections)
    ref = random.choice(range(len(moduleList)))
    child = C_Module(-1, moduleList[ref], ref)
    child.theta = random.randint(0, 3)          # drawn but unused, as in the reference
    rule_module.availableConnections.remove(con)
    child.parentConnectionSite = con
    rule_module.children.append(child)


class Rule:
    def __init__(self, moduleRef, moduleList):
        self.module = C_Module(-1, moduleList[moduleRef], moduleRef)
        self.max_children = len(self.module.availableConnections)
        self.n_children = random.randint(0, self.max_children)
        self.moduleList = moduleList
        self.moduleRef = moduleRef
        for _ in range(self.n_children):
            _random_product(self.module, moduleList)

    def mutate(self, MORPH_MUTATIONRATE, MUTATION_RATE, MUT_SIGMA):
        self.moduleList[self.moduleRef].mutate(MORPH_MUTATIONRATE, MUTATION_RATE, MUT_SIGMA)
        if random.uniform(0.0, 1.0) < MORPH_MUTATIONRATE:
            if self.n_children < self.max_children - 1:
                self.n_children += 1
                _random_product(self.module, self.moduleList)
        if random.uniform(0.0, 1.0) < MORPH_MUTATIONRATE:
            if self.n_children > 0:
                self.n_children -= 1
                victim = random.choice(self.module.children)
                self.module.availableConnections.append(victim.parentConnectionSite)
                self.module.children.remove(victim)

    def update(self, index, share=False):
        """Fresh copies of this rule's products, numbered from index+1 (LSystem.py:114-124)."""
        out = []
        for c in self.module.children:
            index += 1
            if share:                       # field-by-field copy of a product symbol (ints, enum members, a list of enum members)
                sym = C_Module.__new__(C_Module)
                sym.__dict__.update(c.__dict__)
                sym.availableConnections = list(c.availableConnections)
            else:
                sym = copy.deepcopy(c)
            sym.children = []
            sym.index = index
            sym.handled = False
            out.append(sym)
        return index, out


class LSystem:
    def __init__(self, moduleList, config=None):
        self.moduleList = moduleList          # kept by reference, like the reference does
        if config is not None:
            self.treeDepth = int(config['morphology']['max_depth'])
            self.maxModules = int(config['morphology']['max_size'])
        else:
            self.treeDepth = 8
            self.maxModules = 20
        self.rules = [Rule(i, moduleList) for i in range(len(moduleList))]

    def create(self, treedepth, share=False):
        """``share=True`` (used by the population flattener, which only READS the tree): the nodes reference the type's
        module and controller instead of deep copies of them - the same tree, 4x cheaper to build. The default is the
        reference's behaviour (every node owns its copies)."""
        # the argument is ignored; self.treeDepth is used (LSystem.py:157,165)
        if share:
            base = C_Module.__new__(C_Module)
            base.__dict__.update(self.rules[0].module.__dict__)
        else:
            base = copy.deepcopy(self.rules[0].module)
        base.children = []
        base.index = 0
        index = 0
        for _ in range(self.treeDepth):
            index = self.iterate(base, index, 0, share)
        tree = _tree.Tree(self.moduleList)
        self.recursiveNodeGen(-1, base, tree, 0, share)
        return tree

    def iterate(self, currentSymbol, index, depth, share=False):
        if index > self.maxModules:
            return index
        if not currentSymbol.handled:
            currentSymbol.handled = True
            if len(currentSymbol.children) > 0:
                raise Exception("if symbol was not handled it shouldn't contain children")
            index, symbols = self.rules[currentSymbol.moduleRef].update(index, share)
            for s in symbols:
                s.parent = currentSymbol.index
                currentSymbol.children.append(s)
        else:
            for c in currentSymbol.children:
                index = self.iterate(c, index, depth + 1, share)
        return index

    def recursiveNodeGen(self, parentIndex, m, tree, nodeCounter, share=False):
        if nodeCounter > self.maxModules:
            return nodeCounter
        proto = self.moduleList[m.moduleRef]
        node = _tree.Node(m.index, parentIndex, m.moduleRef, m.parentConnectionSite,
                          proto.controller if share else copy.deepcopy(proto.controller))
        node.module_ = proto if share else copy.deepcopy(proto)
        tree.nodes.append(node)
        for c in m.children:
            nodeCounter += 1
            nodeCounter = self.recursiveNodeGen(c.parent, c, tree, nodeCounter, share)
        return nodeCounter

    def mutate(self, MORPH_MUTATIONRATE, MUTATION_RATE, MUT_SIGMA):
        for m in self.moduleList:
            m.mutate(MORPH_MUTATIONRATE, MUTATION_RATE, MUT_SIGMA)
        for r in self.rules:
            r.mutate(MORPH_MUTATIONRATE, MUTATION_RATE, MUT_SIGMA)
