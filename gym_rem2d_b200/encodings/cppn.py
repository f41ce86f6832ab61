"""Feed-forward CPPN genome — stand-in for the reference's neat-python dependency.

The reference wraps ``neat.DefaultGenome`` / ``neat.nn.FeedForwardNetwork`` (NeuralNetwork/NEAT_NN.py:
15-78) configured by NeuralNetwork/config (3 inputs, 10 hidden, ``partial_direct 0.5`` initial wiring,
sum aggregation, 15 activation options, sin default, weights/bias/response in the ranges below).
neat-python 0.92 is a third-party package that is not available offline, so this module restates the
*published* NEAT genome behaviour the wrapper relies on: node genes (bias, response, activation),
connection genes (weight, enabled), structural mutations that keep the graph acyclic, and layered
feed-forward evaluation ``act(bias + response * sum(w_i x_i))``. It is not random-stream compatible
with neat-python; morphologies from this encoding are therefore pinned only
against this implementation.
"""
import random

from .activations import Activation, CPPN_ORDER

_CFG = dict(
    num_hidden=10, connection_fraction=0.5,
    conn_add_prob=0.4, conn_delete_prob=0.1, node_add_prob=0.4, node_delete_prob=0.1,
    activation_default='sin', activation_mutate_rate=0.1,
    bias_init_mean=0.3, bias_init_stdev=0.3, bias_replace_rate=0.3, bias_mutate_rate=0.3,
    bias_mutate_power=0.3, bias_max_value=0.5, bias_min_value=-1.0,
    response_init_mean=0.3, response_init_stdev=0.3, response_replace_rate=0.3,
    response_mutate_rate=0.3, response_mutate_power=0.3, response_max_value=1.0, response_min_value=-1.0,
    weight_init_mean=0.4, weight_init_stdev=0.3, weight_mutate_rate=0.3, weight_replace_rate=0.3,
    weight_mutate_power=0.3, weight_max_value=1.0, weight_min_value=-1.0,
)


def _clip(v, lo, hi):
    return max(lo, min(hi, v))


class _FloatGene:
    """init / mutate rule shared by bias, response and weight attributes."""

    def __init__(self, cfg, prefix):
        self.cfg, self.p = cfg, prefix

    def init(self):
        c, p = self.cfg, self.p
        return _clip(random.gauss(c[p + '_init_mean'], c[p + '_init_stdev']), c[p + '_min_value'], c[p + '_max_value'])

    def mutate(self, v):
        c, p = self.cfg, self.p
        r = random.random()
        if r < c[p + '_mutate_rate']:
            return _clip(v + random.gauss(0.0, c[p + '_mutate_power']), c[p + '_min_value'], c[p + '_max_value'])
        if r < c[p + '_mutate_rate'] + c[p + '_replace_rate']:
            return self.init()
        return v


class FeedForwardNetwork:
    def __init__(self, input_keys, output_keys, node_evals):
        self.input_nodes = input_keys
        self.output_nodes = output_keys
        self.node_evals = node_evals

    def activate(self, inputs):
        values = {k: 0.0 for k in self.output_nodes}
        for k, v in zip(self.input_nodes, inputs):
            values[k] = v
        for node, act, bias, response, links in self.node_evals:
            s = 0.0
            for src, w in links:
                s += values[src] * w
            values[node] = act(bias + response * s)
        return [values[k] for k in self.output_nodes]


class CPPN:
    def __init__(self, n_inputs, n_outputs, t_config=None):
        self.cfg = dict(_CFG)
        self.input_keys = [-(i + 1) for i in range(n_inputs)]
        self.output_keys = list(range(n_outputs))
        if t_config is not None:
            ea = t_config['ea']
            mm, m, s = float(ea['morphmutation_prob']), float(ea['mutation_prob']), float(ea['mutation_sigma'])
            self.cfg.update(conn_add_prob=mm, conn_delete_prob=mm, node_add_prob=mm, node_delete_prob=mm,
                            activation_mutate_rate=m, weight_mutate_power=s, weight_replace_rate=m,
                            weight_mutate_rate=m, bias_replace_rate=mm, bias_mutate_rate=m, bias_mutate_power=s,
                            response_replace_rate=mm, response_mutate_rate=m, response_mutate_power=s)
        self._bias = _FloatGene(self.cfg, 'bias')
        self._resp = _FloatGene(self.cfg, 'response')
        self._w = _FloatGene(self.cfg, 'weight')
        # node genes: key -> [bias, response, activation name]; connection genes: (src, dst) -> [w, enabled]
        self.nodes = {}
        self.conns = {}
        for k in self.output_keys:
            self.nodes[k] = self._new_node()
        self._next_key = n_outputs
        hidden = []
        for _ in range(self.cfg['num_hidden']):
            k = self._next_key
            self._next_key += 1
            self.nodes[k] = self._new_node()
            hidden.append(k)
        # partial_direct: each candidate connection (input->hidden, hidden->output, input->output)
        # is kept with probability connection_fraction
        cand = [(i, h) for i in self.input_keys for h in hidden]
        cand += [(h, o) for h in hidden for o in self.output_keys]
        cand += [(i, o) for i in self.input_keys for o in self.output_keys]
        random.shuffle(cand)
        for c in cand[:int(round(len(cand) * self.cfg['connection_fraction']))]:
            self.conns[c] = [self._w.init(), True]

    def _new_node(self):
        return [self._bias.init(), self._resp.init(), self.cfg['activation_default']]

    # -- structural helpers ------------------------------------------------------------------
    def _creates_cycle(self, src, dst):
        if src == dst:
            return True
        seen = {dst}
        frontier = [dst]
        while frontier:
            n = frontier.pop()
            for (a, b) in self.conns:
                if a == n and b not in seen:
                    if b == src:
                        return True
                    seen.add(b)
                    frontier.append(b)
        return False

    def mutate(self):
        c = self.cfg
        if random.random() < c['node_add_prob'] and self.conns:
            key = random.choice(list(self.conns))
            w, _ = self.conns[key]
            self.conns[key][1] = False
            k = self._next_key
            self._next_key += 1
            self.nodes[k] = self._new_node()
            self.conns[(key[0], k)] = [1.0, True]
            self.conns[(k, key[1])] = [w, True]
        if random.random() < c['node_delete_prob']:
            hidden = [k for k in self.nodes if k not in self.output_keys]
            if hidden:
                k = random.choice(hidden)
                del self.nodes[k]
                for key in [key for key in self.conns if k in key]:
                    del self.conns[key]
        if random.random() < c['conn_add_prob']:
            dst = random.choice(list(self.nodes))
            src = random.choice(list(self.nodes) + self.input_keys)
            if (src, dst) in self.conns:
                self.conns[(src, dst)][1] = True
            elif src not in self.output_keys and not self._creates_cycle(src, dst):
                self.conns[(src, dst)] = [self._w.init(), True]
        if random.random() < c['conn_delete_prob'] and self.conns:
            del self.conns[random.choice(list(self.conns))]
        for gene in self.conns.values():
            gene[0] = self._w.mutate(gene[0])
        for gene in self.nodes.values():
            gene[0] = self._bias.mutate(gene[0])
            gene[1] = self._resp.mutate(gene[1])
            if random.random() < c['activation_mutate_rate']:
                gene[2] = random.choice(CPPN_ORDER)

    def getPhenotype(self):
        """Layered evaluation order over the enabled connections that reach an output."""
        enabled = [k for k, g in self.conns.items() if g[1]]
        required = set(self.output_keys)
        grew = True
        while grew:
            grew = False
            for a, b in enabled:
                if b in required and a not in required and a not in self.input_keys:
                    required.add(a)
                    grew = True
        done = set(self.input_keys)
        evals = []
        pending = set(required)
        while pending:
            ready = sorted(n for n in pending
                           if all(a in done for a, b in enabled if b == n and (a in required or a in self.input_keys)))
            if not ready:
                break
            for n in ready:
                links = [(a, self.conns[(a, n)][0]) for a, b in enabled if b == n and a in done]
                bias, resp, act = self.nodes[n]
                evals.append((n, Activation(act), bias, resp, links))
            done.update(ready)
            pending.difference_update(ready)
        return FeedForwardNetwork(self.input_keys, self.output_keys, evals)

    def update(self, inputs):
        return self.getPhenotype().activate(inputs)
