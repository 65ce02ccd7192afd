"""``function``: a plain interpreter of the miniature graph (see ``__init__``)."""
import numpy as np

from .graph.basic import Constant


def function(inputs, outputs):
    single = not isinstance(outputs, (list, tuple))
    outs = [outputs] if single else list(outputs)

    def run(*values):
        if len(values) != len(inputs):
            raise TypeError('expected %d arguments' % len(inputs))
        memo = {id(v): np.asarray(x, dtype=np.float64) for v, x in zip(inputs, values)}

        def ev(v):
            if id(v) in memo:
                return memo[id(v)]
            if isinstance(v, Constant):
                return v.data
            node = v.owner
            if node is None:
                raise ValueError('no value for input variable %r' % (v,))
            storage = [[None] for _ in node.outputs]
            node.op.perform(node, [ev(i) for i in node.inputs], storage)
            for o, cell in zip(node.outputs, storage):
                memo[id(o)] = cell[0]
            return memo[id(v)]
        res = [ev(o) for o in outs]
        return res[0] if single else res
    return run
