"""Variables and Apply nodes of the miniature stand-in (see ``pytensor/__init__.py``)."""
from types import SimpleNamespace

import numpy as np


class Apply:
    def __init__(self, op, inputs, outputs):
        self.op, self.inputs, self.outputs = op, list(inputs), list(outputs)
        for i, out in enumerate(self.outputs):
            out.owner, out.index = self, i


class Variable:
    def __init__(self, type, name=None):
        self.type, self.name = type, name
        self.owner, self.index = None, 0
        self.tag = SimpleNamespace()

    ndim = property(lambda self: self.type.ndim)
    dtype = property(lambda self: self.type.dtype)

    def __repr__(self):
        return self.name or '<%s>' % self.type

    # the operations the wrapper and the reference's test use
    def reshape(self, shape):
        from .. import tensor
        return tensor.reshape(self, shape)

    def sum(self, axis=None):
        from .. import tensor
        return tensor.sum(self, axis)

    def __getitem__(self, idx):
        from .. import tensor
        return tensor.subtensor(self, idx)

    def __mul__(self, other):
        from .. import tensor
        return tensor.mul(self, other)

    __rmul__ = __mul__

    def __add__(self, other):
        from .. import tensor
        return tensor.add(self, other)

    __radd__ = __add__

    def __neg__(self):
        from .. import tensor
        return tensor.neg(self)


class Constant(Variable):
    def __init__(self, type, data, name=None):
        super().__init__(type, name)
        self.data = np.asarray(data)
