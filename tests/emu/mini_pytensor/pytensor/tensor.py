"""Tensor types and the few operations needed (miniature stand-in, see ``__init__``)."""
import numpy as np

from .graph.basic import Apply, Constant, Variable
from .graph.op import Op


class TensorType:
    def __init__(self, dtype, shape):
        self.dtype, self.shape = str(np.dtype(dtype)), tuple(shape)

    ndim = property(lambda self: len(self.shape))

    def __call__(self, name=None):
        return Variable(self, name)

    def __repr__(self):
        return 'TensorType(%s, %s)' % (self.dtype, self.shape)


dscalar = TensorType('float64', ())
dvector = TensorType('float64', (None,))
dmatrix = TensorType('float64', (None, None))
dtensor3 = TensorType('float64', (None, None, None))


def as_tensor_variable(x, dtype=None, name=None):
    if isinstance(x, Variable):
        if dtype is not None and x.dtype != str(np.dtype(dtype)):
            raise TypeError('cannot cast a %s variable to %s' % (x.dtype, dtype))
        return x
    data = np.asarray(x, dtype=dtype if dtype is not None else None)
    if data.dtype.kind in 'iub' and dtype is None:
        data = data.astype(np.float64)
    return Constant(TensorType(data.dtype, data.shape), data, name)


class Lambda(Op):
    """An operation given by a numpy function of its inputs; ``grad_fn(inputs, output, g)`` returns
    the gradient expressions (built from further Lambda nodes)."""

    def __init__(self, name, fn, ndim_out, grad_fn=None):
        self.name, self.fn, self.ndim_out, self.grad_fn = name, fn, ndim_out, grad_fn

    def make_node(self, *inputs):
        ins = [as_tensor_variable(x) for x in inputs]
        return Apply(self, ins, [TensorType('float64', (None,) * self.ndim_out)()])

    def perform(self, node, inputs, output_storage):
        output_storage[0][0] = np.asarray(self.fn(*inputs), dtype=np.float64)

    def grad(self, inputs, output_grads):
        if self.grad_fn is None:
            raise NotImplementedError('no gradient for ' + self.name)
        return self.grad_fn(inputs, output_grads[0])


def _like(g, x):
    """g summed / reshaped to the shape of x (undoes numpy broadcasting)."""
    def fn(gv, xv):
        gv = np.asarray(gv)
        while gv.ndim > xv.ndim:
            gv = gv.sum(axis=0)
        for ax, n in enumerate(xv.shape):
            if n == 1 and gv.shape[ax] != 1:
                gv = gv.sum(axis=ax, keepdims=True)
        return gv.reshape(xv.shape)
    return Lambda('sum_to_shape', fn, x.ndim)(g, x)


def reshape(x, shape):
    shape = tuple(shape)
    return Lambda('reshape', lambda v: np.reshape(v, shape), len(shape),
                  lambda ins, g: [Lambda('reshape_like', lambda gv, xv: np.reshape(gv, xv.shape), ins[0].ndim)(g, ins[0])])(x)


def subtensor(x, idx):
    x = as_tensor_variable(x)
    ndim = np.empty((2,) * x.ndim)[idx].ndim

    def back(gv, xv):
        z = np.zeros_like(xv)
        np.add.at(z, idx, gv)
        return z
    return Lambda('subtensor', lambda v: v[idx], ndim,
                  lambda ins, g: [Lambda('subtensor_grad', back, ins[0].ndim)(g, ins[0])])(x)


def mul(a, b):
    a, b = as_tensor_variable(a), as_tensor_variable(b)
    return Lambda('mul', lambda u, v: u * v, max(a.ndim, b.ndim),
                  lambda ins, g: [_like(mul(g, ins[1]), ins[0]), _like(mul(g, ins[0]), ins[1])])(a, b)


def add(a, b):
    a, b = as_tensor_variable(a), as_tensor_variable(b)
    return Lambda('add', lambda u, v: u + v, max(a.ndim, b.ndim),
                  lambda ins, g: [_like(g, ins[0]), _like(g, ins[1])])(a, b)


def neg(a):
    return Lambda('neg', lambda u: -u, as_tensor_variable(a).ndim, lambda ins, g: [neg(g)])(a)


def sum(x, axis=None):  # noqa: A001
    x = as_tensor_variable(x)
    if axis is None:
        axes = tuple(range(x.ndim))
    else:
        axes = tuple(a % x.ndim for a in (axis if isinstance(axis, (tuple, list)) else (axis,)))

    def back(gv, xv):
        return np.broadcast_to(np.expand_dims(np.asarray(gv), axes), xv.shape).copy()
    return Lambda('sum', lambda v: np.sum(v, axis=axes), x.ndim - len(axes),
                  lambda ins, g: [Lambda('sum_grad', back, ins[0].ndim)(g, ins[0])])(x)


def zeros_like(x):
    x = as_tensor_variable(x)
    return Lambda('zeros_like', lambda v: np.zeros_like(v), x.ndim)(x)


def concatenate(parts, axis=0):
    parts = [as_tensor_variable(p) for p in parts]

    def back_k(k):
        def fn(gv, *xs):
            sizes = [x.shape[axis] for x in xs]
            start = int(np.sum(sizes[:k]))
            return np.take(gv, range(start, start + sizes[k]), axis=axis)
        return fn
    return Lambda('concatenate', lambda *vs: np.concatenate(vs, axis=axis), parts[0].ndim,
                  lambda ins, g: [Lambda('split', back_k(k), ins[k].ndim)(g, *ins) for k in range(len(ins))])(*parts)


def linspace(start, stop, num):
    """A symbolic (non-constant) vector, as in PyTensor, so that gradients with respect to it exist."""
    return Lambda('linspace', lambda: np.linspace(float(start), float(stop), int(num)), 1, lambda ins, g: [])()


def grad(cost, wrt):
    from .gradient import grad as _grad
    return _grad(cost, wrt)
