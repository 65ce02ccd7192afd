"""Reverse-mode differentiation over the miniature graph (see ``__init__``)."""
from .graph.basic import Constant, Variable


class _Special(Variable):
    def __init__(self, text):
        super().__init__(None)
        self.text = text

    def __repr__(self):
        return self.text

    __str__ = __repr__


def disconnected():
    return _Special('<DisconnectedType>')


def grad_not_implemented(op, x_pos, x, comment=''):
    return _Special('<grad not implemented: %s input %d>' % (type(op).__name__, x_pos))


def _topo(cost):
    order, seen = [], set()

    def visit(v):
        node = v.owner
        if node is None or id(node) in seen:
            return
        seen.add(id(node))
        for i in node.inputs:
            visit(i)
        order.append(node)
    visit(cost)
    return order


def grad(cost, wrt):
    from . import tensor
    single = not isinstance(wrt, (list, tuple))
    wrts = [wrt] if single else list(wrt)
    grads = {id(cost): tensor.as_tensor_variable(1.0)}
    for node in reversed(_topo(cost)):
        outs = [grads.get(id(o)) for o in node.outputs]
        if all(g is None for g in outs):
            continue
        outs = [g if g is not None else disconnected() for g in outs]
        in_grads = node.op.grad(node.inputs, outs)
        if len(in_grads) != len(node.inputs):
            raise ValueError('%s.grad returned %d gradients for %d inputs'
                             % (type(node.op).__name__, len(in_grads), len(node.inputs)))
        for v, g in zip(node.inputs, in_grads):
            if isinstance(g, _Special) or isinstance(v, Constant):
                if isinstance(g, _Special) and 'not implemented' in g.text and any(v is w for w in wrts):
                    raise NotImplementedError(g.text)
                continue
            grads[id(v)] = g if id(v) not in grads else tensor.add(grads[id(v)], g)
    result = []
    for w in wrts:
        if id(w) not in grads:
            raise ValueError('cost does not depend on %r' % (w,))
        result.append(grads[id(w)])
    return result[0] if single else result
