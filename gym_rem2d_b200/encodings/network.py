"""Network-driven growth encoding (reference: Encodings/Network_Encoding.py:27-222).

A network (CPPN or the cell network grown by the cellular encoding) is queried once per free
connection site with (normalised depth, parent module type, site side) and answers: grow a module
here?, which type, its morphology parameters and its four sine-controller parameters. The network is
only evaluated at expansion time — at run time the creature is driven by plain sine controllers.
"""
import copy
from enum import Enum

from .. import tree as _tree
from . import cellular
from . import cppn as _cppn

MAX_MODULES = 20


class NETWORK_TYPE(Enum):          # Network_Encoding.py:18-20 (kept as an Enum so that pickles are interchangeable)
    CPPN = 0
    CE = 1


class C_Module:
    def __init__(self, index, module, moduleRef):
        self.index = index
        self.parent = -1
        self.moduleRef = moduleRef
        self.availableConnections = copy.deepcopy(module.available)
        self.children = []
        self.theta = -1
        self.parentConnectionSite = None
        self.handled = False
        self.module = copy.deepcopy(module)
        self.controller = None


class NN_enc:
    def __init__(self, modulelist, type, config=None):
        self.moduleList = copy.deepcopy(modulelist)
        n_inputs, n_outputs = 3, 10
        self.outputs = [0] * n_outputs
        self.inputs = []
        if config is not None:
            self.maxTreeDepth = int(config['morphology']['max_depth'])
            self.maxModules = int(config['morphology']['max_size'])
        else:
            self.maxTreeDepth = 7
            self.maxModules = 20
        if type not in ("CPPN", "CE"):
            raise Exception("Cannot create network encoding, unknown network type %r" % (type,))
        self.networkType = NETWORK_TYPE.CPPN if type == "CPPN" else NETWORK_TYPE.CE
        if type == "CPPN":
            self.nn_g = _cppn.CPPN(n_inputs, n_outputs, t_config=config)
        elif type == "CE":
            self.nn_g = cellular.CE(config=config)
            self.nn_g.mutate(0.5, 0.5, 0.5)
            self.nn_g.create()
        for mod in self.moduleList:
            mod.mutate(0.5, 0.5, 0.5)

    def _query(self, inputs):
        if self.networkType == NETWORK_TYPE.CPPN:
            return list(self.nn_p.activate(inputs))
        if self.networkType == NETWORK_TYPE.CE:
            return self.nn_p.update(inputs, requested_number_of_outputs=9)
        raise Exception("Cannot update network, no network type found")

    def update(self, index, par_symb, depth):
        """Ask the network about every free site of ``par_symb`` (Network_Encoding.py:86-139)."""
        new_symbols = []
        if depth > self.maxTreeDepth or index > self.maxModules:
            return index, new_symbols
        n_types = len(self.moduleList)
        for con in par_symb.availableConnections:
            out = self._query([
                float(1) - (float(2) * (float(depth) / float(self.maxTreeDepth))),
                float(1) - (float(2) * (float(par_symb.moduleRef + 1) / float(n_types))),
                con.value[0],
            ])
            if not out[0] > 0.5:
                continue
            out[1] = max(-1., min(1., out[1])) if abs(out[1]) > 1. else out[1]
            ref = int(((out[1] * 0.5) + 0.5) * float(n_types - 1))
            ref = min(max(ref, 0), n_types - 1)
            proto = self.moduleList[ref]
            sym = C_Module(index, proto, ref)
            sym.module.setMorph(out[2], out[3], out[4])
            ctl = copy.deepcopy(proto.controller)
            ctl.setControl(out[5], out[6], out[7], out[8], proto.angle)
            sym.controller = ctl
            sym.parent = par_symb.index
            sym.parentConnectionSite = con
            par_symb.children.append(sym)
            new_symbols.append(sym)
            index += 1
        return index, new_symbols

    def mutate(self, MORPH_MUTATION_RATE, MUTATION_RATE, MUT_SIGMA, TREE_DEPTH=None):
        if self.networkType == NETWORK_TYPE.CPPN:
            self.nn_g.mutate()
        elif self.networkType == NETWORK_TYPE.CE:
            self.nn_g.mutate(MORPH_MUTATION_RATE, MUTATION_RATE, MUT_SIGMA)
            self.nn_g.create()
        for mod in self.moduleList:
            mod.mutate(MORPH_MUTATION_RATE, MUTATION_RATE, MUT_SIGMA)

    def iterate(self, current_symbol, index, depth):
        if not current_symbol.handled:
            current_symbol.handled = True
            if len(current_symbol.children) > 0:
                raise Exception("if symbol was not handled it shouldn't contain children")
            index, symbols = self.update(index, current_symbol, depth)
            for s in symbols:
                s.parent = current_symbol.index
        else:
            for c in current_symbol.children:
                index = self.iterate(c, index, depth + 1)
        return index

    def create(self, treedepth):
        self.maxTreeDepth = treedepth
        if self.networkType == NETWORK_TYPE.CE:
            self.nn_g.create()
            self.nn_p = self.nn_g
        elif self.networkType == NETWORK_TYPE.CPPN:
            self.nn_p = self.nn_g.getPhenotype()
        base = C_Module(0, self.moduleList[0], -1)      # the root's type is -1 (Network_Encoding.py:189)
        base.controller = copy.deepcopy(self.moduleList[0].controller)
        index = 1
        for _ in range(treedepth):
            index = self.iterate(base, index, 0)
        self.nn_p = None
        tree = _tree.Tree(self.moduleList)
        self.recursiveNodeGen(-1, base, tree, 0)
        return tree

    def recursiveNodeGen(self, parentIndex, m, tree, nodeCounter):
        if nodeCounter > MAX_MODULES:
            return nodeCounter
        node = _tree.Node(m.index, parentIndex, m.moduleRef, m.parentConnectionSite, m.controller)
        node.module_ = m.module
        tree.nodes.append(node)
        for c in m.children:
            nodeCounter += 1
            nodeCounter = self.recursiveNodeGen(c.parent, c, tree, nodeCounter)
        return nodeCounter
