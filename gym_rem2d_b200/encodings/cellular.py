"""Cellular encoding: a short list of division schemes grows a small feed-forward cell network.

Reference: Encodings/Cellular_Encoding.py:52-335 (Scheme, Cell, Link, CE). Only the parts that the
network encoding exercises (construct, mutate, create, update) are restated; plotting is not.
The implementation deliberately keeps the reference's observable quirks (noted inline) because the
expanded morphologies are pinned against the reference's own output.
"""
import random

import numpy as np

from .activations import Activation, CE_ORDER

N_ITERATIONS = 10
N_SCHEMES = 10
MAX_CELLS = 50
_WIDTH = 1000
_HEIGHT = 1000
_DIS = 100


class Scheme:
    def __init__(self, n_schemes):
        self.p_symbols = []
        self.weights = []
        self.thresholds = []
        self.activationFunctions = []
        for _ in range(random.randint(1, 2)):
            self.p_symbols.append(random.randint(0, n_schemes - 1))
            self.weights.append(random.uniform(-1, 1))
            self.thresholds.append(random.uniform(0.5, 1))
            self.activationFunctions.append(Activation(random.choice(CE_ORDER)))
        self.type = random.randint(0, 1)          # 0 sequential division, 1 parallel division


class Link:
    def __init__(self, weight, t_index):
        self.weight = weight
        self.c_index = t_index


class Cell:
    def __init__(self, type, pos, threshold, index, activationFunction):
        self.type = type
        self.pos = pos
        self.input_indices = []
        self.index = index
        self.activationFunction = activationFunction
        self.output_links = []
        self.activity = 0.0
        self.layer = 0
        self.threshold = threshold

    def update(self):
        self.activity = max(-1.0, min(1.0, self.activity)) if abs(self.activity) > 1.0 else self.activity
        total = 0.0
        for link in self.output_links:
            total += self.activationFunction(self.activity * link.weight)
        return total


class CE:
    def __init__(self, config=None):
        self.schemes = [Scheme(N_SCHEMES) for _ in range(N_SCHEMES)]
        self.init()
        self.index = 0

    def init(self):
        """input cell -> one seed cell -> output cell (Cellular_Encoding.py:114-139)."""
        s0 = self.schemes[0]
        self.cells = []
        i_pos = [_WIDTH * 0.5, 10]
        o_pos = [_WIDTH * 0.5, _HEIGHT - 10]
        self.inputCell = Cell(-1, i_pos, -1, 0, s0.activationFunctions[0])
        self.outputCell = Cell(-1, o_pos, -1, 1, s0.activationFunctions[0])
        seed = Cell(0, [(i_pos[0] + o_pos[0]) * 0.5, (i_pos[1] + o_pos[1]) * 0.5], s0.thresholds[0], 2,
                    s0.activationFunctions[0])
        self.index = 3
        seed.output_links.append(Link(1.0, self.outputCell.index))
        seed.input_indices.append(self.inputCell.index)
        seed.layer = 1
        self.cells.append(seed)
        self.inputCell.output_links.append(Link(1.0, seed.index))
        self.outputCell.input_indices.append(seed.index)

    def create(self):
        self.init()
        for i in range(N_ITERATIONS):
            self.iterate(i + 1)

    def mutate(self, MORPH_MUTATION_RATE, MUTATION_RATE, MUT_SIGMA):
        for scheme in self.schemes:
            if random.uniform(0, 1) < MORPH_MUTATION_RATE:
                scheme.p_symbols = []
                n_symbols = random.randint(1, 2)
                scheme.weights = []
                for _ in range(n_symbols):
                    scheme.p_symbols.append(random.randint(0, N_SCHEMES - 1))
                    scheme.weights.append(random.uniform(0, 1))
                    # thresholds / activation lists are appended to, not reset (reference quirk)
                    scheme.thresholds.append(random.uniform(-1, 1))
                    scheme.activationFunctions.append(Activation(random.choice(CE_ORDER)))
            # the next two loops only consume random numbers in the reference (they rebind a loop
            # variable), so they do here too
            for _ in scheme.activationFunctions:
                if random.uniform(0, 1) < MUTATION_RATE:
                    random.choice(CE_ORDER)
            for _ in scheme.p_symbols:
                if random.uniform(0, 1) < MORPH_MUTATION_RATE:
                    random.randint(0, N_SCHEMES - 1)
            for values in (scheme.weights, scheme.thresholds):
                for i, w in enumerate(values):
                    if random.uniform(0, 1) < MUTATION_RATE:
                        w += random.gauss(w, MUT_SIGMA)
                        values[i] = max(-1.0, min(1.0, w)) if abs(w) > 1.0 else w

    def _cell_by_index(self, idx):
        """Last cell in ``self.cells`` with that index (the reference scans without break)."""
        found = None
        for cell in self.cells:
            if cell.index == idx:
                found = cell
        return found

    def update(self, inputs, requested_number_of_outputs=1):
        """Feed ``inputs`` through the grown network (Cellular_Encoding.py:199-259).

        Returns *at least* the requested number of outputs: every pass over the layers appends one
        value per link into the output cell, and passes repeat (activity accumulating) until enough
        values exist.
        """
        for cell in self.cells:
            cell.activity = 0.0
        in_links = self.inputCell.output_links
        for i, link in enumerate(in_links):
            link.weight = inputs[int(float(i) / float(len(in_links)) * 3.0)]
        self.inputCell.activity = 1.0
        drive = self.inputCell.update()
        for link in in_links:
            target = self._cell_by_index(link.c_index)
            if self.outputCell.index == link.c_index:
                target = self.outputCell
            target.activity += drive

        output = []
        n_layers = -1
        for cell in self.cells:
            if cell.layer >= n_layers:
                n_layers = cell.layer
        while len(output) < requested_number_of_outputs:
            for layer in range(n_layers + 1):
                for cell in self.cells:
                    if cell.layer != layer:
                        continue
                    out = cell.update()
                    for link in cell.output_links:
                        if link.c_index == self.outputCell.index:
                            output.append(out)
                        for other in self.cells:
                            if link.c_index == other.index:
                                other.activity += out
        return output

    def iterate(self, iterationNumber):
        n_cells = len(self.cells)
        for i in range(n_cells):
            cell = self.cells[i]
            if len(self.cells) > MAX_CELLS:
                return
            if cell.type == -1:
                continue
            s = self.schemes[cell.type]
            product = s.p_symbols
            if len(product) == 1:
                cell.type = product[0]
                for link in cell.output_links:
                    link.weight = s.weights[0]
                continue
            step = _DIS / np.sqrt(iterationNumber)
            c1 = cell
            if s.type == 0:
                # sequential division: c2 is inserted between c1 and c1's targets
                p_c1 = [cell.pos[0], cell.pos[1] - step]
                p_c2 = [cell.pos[0], cell.pos[1] + step]
                c1.type = product[0]
                c1.pos = p_c1
                c1.threshold = s.thresholds[0]
                c1.activationFunction = s.activationFunctions[0]
                c2 = Cell(product[1], p_c2, s.thresholds[1], self.index, s.activationFunctions[1])
                self.index += 1
                c2.layer = c1.layer
                c2.input_indices.append(c1.index)
                # live-iteration + remove: every other link is moved (reference quirk)
                for out in c1.output_links:
                    out.weight = s.weights[1]
                    c2.output_links.append(out)
                    c1.output_links.remove(out)
                c1.output_links.append(Link(s.weights[0], c2.index))
                c2.layer += 1
                self.cells.append(c2)
            elif s.type == 1:
                # parallel division: c2 duplicates c1's inputs and outputs
                p_c1 = [cell.pos[0] - step, cell.pos[1]]
                p_c2 = [cell.pos[0] + step, cell.pos[1]]
                c1.type = product[0]
                c1.pos = p_c1
                c1.threshold = s.thresholds[0]
                c1.activationFunction = s.activationFunctions[0]
                c2 = Cell(product[1], p_c2, s.thresholds[1], self.index, s.activationFunctions[1])
                self.index += 1
                c2.layer = c1.layer
                for out in c1.output_links:
                    t_c = self._cell_by_index(out.c_index)
                    if out.c_index == self.outputCell.index:
                        t_c = self.outputCell
                    if t_c is None:
                        raise Exception("Target cell is none")
                    c2.output_links.append(Link(s.weights[1], t_c.index))
                    t_c.input_indices.append(c2.index)
                    out.weight = s.weights[0]
                for inp in c1.input_indices:
                    c2.input_indices.append(inp)
                    src = self._cell_by_index(inp)
                    if self.inputCell.index == inp:
                        src = self.inputCell
                    elif self.outputCell.index == inp:
                        src = self.inputCell
                    src.output_links.append(Link(src.output_links[0].weight, c2.index))
                self.cells.append(c2)
