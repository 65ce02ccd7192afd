"""A MINIATURE stand-in for PyTensor -- TEST INFRASTRUCTURE ONLY.

PyTensor is not part of the build image, so the graph half of
``sunode_b200/wrappers/as_pytensor.py`` (``solve_ivp``, ``Op.make_node`` through ``itypes`` /
``otypes``, ``Op.grad``) could not be executed at all.  This package implements just the slice of
the PyTensor API that code and the reference's own test (``sunode/test_pytensor.py``) touch --
symbolic variables, ``Op`` / ``Apply``, a handful of tensor operations with their gradients,
reverse-mode ``grad`` and an interpreting ``function`` -- so that the wrapper's graph code RUNS in
the CPU test tier (on the stand-in CUDA driver).  It makes no attempt to be PyTensor: no
optimisation, no C code, no broadcasting patterns, first derivatives only.  It is put on the path
by ``tests/emu/pytensor_graph_child.py`` and by nothing else.
"""
from . import gradient, tensor  # noqa: F401
from .compile import function  # noqa: F401
from .gradient import grad  # noqa: F401
