"""Location of build products that must travel with the source tree (cubins, host .so files).

They are kept in-tree (``sunode_b200/_cache``) so that objects compiled ahead of time by
``__graft_entry__.build()`` are found again on the GPU box; ``SUNODE_B200_CACHE`` overrides it,
and an unwritable tree falls back to a per-user temp directory.
"""
from __future__ import annotations

import os
import tempfile

_HERE = os.path.dirname(os.path.abspath(__file__))


def cache_dir() -> str:
    path = os.environ.get('SUNODE_B200_CACHE') or os.path.join(_HERE, '_cache')
    try:
        os.makedirs(path, exist_ok=True)
        probe = os.path.join(path, '.w%d' % os.getpid())
        with open(probe, 'w'):
            pass
        os.unlink(probe)
        return path
    except OSError:
        path = os.path.join(tempfile.gettempdir(), 'sunode_b200_cache_%d' % os.getuid())
        os.makedirs(path, exist_ok=True)
        return path


def atomic_write(path: str, data: bytes) -> None:
    tmp = '%s.tmp%d' % (path, os.getpid())
    with open(tmp, 'wb') as fh:
        fh.write(data)
    os.replace(tmp, path)
