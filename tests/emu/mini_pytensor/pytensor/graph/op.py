"""``Op`` of the miniature stand-in: ``itypes`` / ``otypes`` drive ``make_node`` as in PyTensor."""
from .basic import Apply


class Op:
    itypes = None
    otypes = None

    def make_node(self, *inputs):
        from ..tensor import as_tensor_variable
        if self.itypes is None or self.otypes is None:
            raise NotImplementedError('make_node needs itypes / otypes')
        if len(inputs) != len(self.itypes):
            raise ValueError('%s takes %d inputs, got %d' % (type(self).__name__, len(self.itypes), len(inputs)))
        ins = []
        for x, t in zip(inputs, self.itypes):
            v = as_tensor_variable(x)
            if v.type.ndim != t.ndim or v.type.dtype != t.dtype:
                raise TypeError('%s: expected %s, got %s' % (type(self).__name__, t, v.type))
            ins.append(v)
        return Apply(self, ins, [t() for t in self.otypes])

    def __call__(self, *inputs):
        node = self.make_node(*inputs)
        return node.outputs[0] if len(node.outputs) == 1 else list(node.outputs)

    def perform(self, node, inputs, output_storage):
        raise NotImplementedError

    def grad(self, inputs, output_grads):
        raise NotImplementedError('%s has no gradient' % type(self).__name__)
