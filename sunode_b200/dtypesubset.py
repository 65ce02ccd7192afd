"""Named / nested variable layouts and their flat float64 index maps.

Behavioural counterpart of the reference's ``sunode/dtypesubset.py`` (DTypeSubset,
``as_flattened``/``as_nested``, reference file lines 10-33 and 71-288): a (nested) dict
``name -> shape`` is turned into a numpy structured dtype plus the flat slices that the
integrator kernels index with.  The flat order is declaration order (depth first), which is also
the memory order of the structured dtype, so ``record.view(float64)`` and the flat vector the
CUDA kernels consume are the same bytes.

The implementation is organised around a flat list of ``_Leaf`` records (one per array-valued
entry) from which all dtypes are derived, rather than the reference's single recursive
constructor.
"""
from __future__ import annotations

import dataclasses
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

try:  # pandas is only used to normalise user supplied coordinate labels
    import pandas as pd
except Exception:  # pragma: no cover - pandas is present in the target image
    pd = None  # type: ignore

Path = Tuple[str, ...]
Shape = Tuple[int, ...]


# --------------------------------------------------------------------------------------
# dict helpers (reference: dtypesubset.py:10-64)
# --------------------------------------------------------------------------------------

def as_flattened(vals: Dict[str, Any], base: Optional[Path] = None) -> Dict[Path, Any]:
    """``{'a': {'b': 1}} -> {('a', 'b'): 1}`` (depth-first, insertion ordered)."""
    prefix: Path = tuple(base) if base else ()
    flat: Dict[Path, Any] = {}

    def walk(pre: Path, node: Dict[str, Any]) -> None:
        for key, item in node.items():
            if isinstance(item, dict):
                walk(pre + (key,), item)
            else:
                flat[pre + (key,)] = item

    walk(prefix, vals)
    return flat


def as_nested(vals: Dict[Path, Any]) -> Dict[str, Any]:
    """Inverse of :func:`as_flattened`."""
    root: Dict[str, Any] = {}
    for path, item in vals.items():
        if len(path) == 0:
            raise ValueError("Empty path.")
        node = root
        for key in path[:-1]:
            node = node.setdefault(key, {})
        if path[-1] in node:
            raise ValueError("Duplicate path %s" % (path,))
        node[path[-1]] = item
    return root


def count_items(dtype: np.dtype) -> int:
    """Number of scalar items in a (possibly nested, possibly sub-array) dtype."""
    if dtype.fields is None:
        return int(np.prod(dtype.shape, dtype=np.int64)) if dtype.shape else 1
    return sum(count_items(sub) for sub, *_ in dtype.fields.values())


def _as_dict(data: np.ndarray) -> Any:
    if data.dtype.fields is None:
        return data
    return {name: _as_dict(data[name]) for name in data.dtype.fields}


def _from_dict(data: np.ndarray, vals: Any) -> None:
    if data.dtype.fields is None:
        data[...] = vals
        return
    for name, (sub, *_rest) in data.dtype.fields.items():
        if sub.fields is not None:
            _from_dict(data[name], vals[name])
        else:
            data[name] = vals[name]


# --------------------------------------------------------------------------------------
# layout
# --------------------------------------------------------------------------------------

@dataclasses.dataclass
class _Leaf:
    path: Path
    dtype: Any
    shape: Shape
    dim_names: List[str]
    flat_start: int
    in_subset: bool

    @property
    def size(self) -> int:
        n = 1
        for length in self.shape:
            n *= length
        return n


def _make_index(values: Any, name: Optional[str] = None) -> Any:
    if pd is not None:
        idx = pd.Index(values)
        if name is not None:
            idx = idx.rename(name)
        return idx
    return list(values)


def _range_index(length: int, name: str) -> Any:
    if pd is not None:
        return pd.RangeIndex(length, name=name)
    return list(range(length))


class DTypeSubset:
    """Layout of a nested set of named arrays, with a distinguished subset of entries.

    Parameters mirror the reference constructor (dtypesubset.py:90-97): ``dims`` maps names to a
    shape (tuple of ints and/or coordinate names, a bare int/str, or a nested dict);
    ``subset_paths`` lists the entries that belong to the subset (for parameters: the ones we
    differentiate with respect to).  With ``fixed_dtype=None`` each leaf is ``(dtype, shape)``.

    Attributes
    ----------
    dtype : structured dtype of the full set
    subset_dtype : packed structured dtype of only the subset entries
    subset_view_dtype : dtype with explicit offsets that views the subset *inside* a ``dtype``
        record (reference dtypesubset.py:185-190), used to scatter subset values in place
    paths, subset_paths : declaration-ordered paths
    flat_slices, flat_shapes : path -> slice into / shape within the flat float vector
    """

    def __init__(
        self,
        dims: Dict[str, Any],
        subset_paths: Sequence[Path],
        fixed_dtype: Optional[Any] = None,
        coords: Optional[Dict[str, Any]] = None,
        dim_basename: str = '',
    ) -> None:
        if coords is None:
            coords = {}
        else:
            coords = {name: _make_index(c) for name, c in coords.items()}
        self._fixed_dtype = fixed_dtype
        self._input_dims = dims
        wanted = {tuple(p) for p in subset_paths}

        leaves: List[_Leaf] = []
        dims_out: Dict[str, Any] = {}
        counter = [0]

        def build(node: Dict[str, Any], prefix: Path, basename: str, dims_node: Dict[str, Any]):
            """Returns (full fields, subset fields, view names, view formats, view offsets, size)."""
            fields: List[Tuple[str, Any, Shape]] = []
            sub_fields: List[Tuple[str, Any, Shape]] = []
            v_names: List[str] = []
            v_formats: List[Any] = []
            v_offsets: List[int] = []
            byte_offset = 0
            for name, val in node.items():
                if isinstance(val, dict):
                    child_dims: Dict[str, Any] = {}
                    # the reference names auto-dimensions of nested entries with this literal
                    # prefix (dtypesubset.py:121); kept so coordinate names are identical
                    child = build(val, prefix + (name,), "dim_basename_%s" % name, child_dims)
                    c_fields, c_sub, c_vn, c_vf, c_vo, c_size = child
                    child_dtype = np.dtype(c_fields)
                    fields.append((name, child_dtype, ()))
                    dims_node[name] = child_dims
                    child_sub_dtype = np.dtype(c_sub)
                    if child_sub_dtype.itemsize > 0:
                        sub_fields.append((name, child_sub_dtype, ()))
                        v_names.append(name)
                        v_formats.append(np.dtype({
                            'names': c_vn, 'formats': c_vf, 'offsets': c_vo,
                            'itemsize': child_dtype.itemsize,
                        }))
                        v_offsets.append(byte_offset)
                    byte_offset += child_dtype.itemsize
                    continue

                if fixed_dtype is None:
                    leaf_dtype, val = val
                else:
                    leaf_dtype = fixed_dtype
                if isinstance(val, (int, str, np.integer)):
                    val = (val,)
                shape: List[int] = []
                dim_names: List[str] = []
                for axis, dim in enumerate(val):
                    if isinstance(dim, str):
                        if dim not in coords:
                            raise KeyError('Unknown dimension name: %s' % dim)
                        shape.append(len(coords[dim]))
                        dim_names.append(dim)
                    else:
                        length = int(dim)
                        auto = '%s_%s_dim%s__' % (basename, name, axis)
                        if auto in coords:
                            raise ValueError(
                                "Can not create two different dimensions with the same name: "
                                "%s." % auto)
                        coords[auto] = _range_index(length, auto)
                        shape.append(length)
                        dim_names.append(auto)
                path = prefix + (name,)
                leaf = _Leaf(path, leaf_dtype, tuple(shape), dim_names, counter[0], path in wanted)
                counter[0] += leaf.size
                leaves.append(leaf)
                dims_node[name] = (leaf_dtype, dim_names)
                field = (name, leaf_dtype, tuple(shape))
                fields.append(field)
                if leaf.in_subset:
                    sub_fields.append(field)
                    v_names.append(name)
                    v_formats.append((leaf_dtype, tuple(shape)))
                    v_offsets.append(byte_offset)
                byte_offset += np.dtype([field]).itemsize
            return fields, sub_fields, v_names, v_formats, v_offsets, byte_offset

        fields, sub_fields, v_names, v_formats, v_offsets, _ = build(dims, (), dim_basename, dims_out)

        self.dtype = np.dtype(fields)
        self.subset_dtype = np.dtype(sub_fields)
        self.subset_view_dtype = np.dtype({
            'names': v_names,
            'formats': v_formats,
            'offsets': v_offsets,
            'itemsize': self.dtype.itemsize,
        })
        self._leaves = leaves
        self.item_count = counter[0]
        self.coords = coords
        self.dims = dims_out
        self.paths = [leaf.path for leaf in leaves]
        # declaration order wins over the order the caller listed the subset in
        self.subset_paths = [leaf.path for leaf in leaves if leaf.in_subset]
        self.flat_slices = {
            leaf.path: slice(leaf.flat_start, leaf.flat_start + leaf.size) for leaf in leaves}
        self.flat_shapes = {leaf.path: leaf.shape for leaf in leaves}
        self._remainder: Optional['DTypeSubset'] = None

    # ------------------------------------------------------------------ counts / index maps
    @property
    def n_subset(self) -> int:
        return count_items(self.subset_dtype)

    @property
    def n_items(self) -> int:
        return count_items(self.dtype)

    @property
    def subset_flat_index(self) -> np.ndarray:
        """Flat positions (into the full vector) of the subset items, in subset order."""
        idx: List[int] = []
        for leaf in self._leaves:
            if leaf.in_subset:
                idx.extend(range(leaf.flat_start, leaf.flat_start + leaf.size))
        return np.asarray(idx, dtype=np.int64)

    @property
    def remainder_flat_index(self) -> np.ndarray:
        """Flat positions of the items *not* in the subset, in declaration order."""
        idx: List[int] = []
        for leaf in self._leaves:
            if not leaf.in_subset:
                idx.extend(range(leaf.flat_start, leaf.flat_start + leaf.size))
        return np.asarray(idx, dtype=np.int64)

    # ------------------------------------------------------------------ conversions
    def set_from_subset(self, value_buffer: np.ndarray, subset_buffer: np.ndarray) -> None:
        value_buffer.view(self.subset_dtype).fill(subset_buffer)

    def as_dataclass(
        self,
        dataclass_name: str,
        flat_subset: Sequence[Any],
        flat_remainder: Sequence[Any],
        item_map: Optional[Callable[[np.ndarray], Any]] = None,
    ) -> Any:
        """Attribute tree (nested dataclasses) whose leaves are taken, in order, from the two
        flat sequences (reference dtypesubset.py:215-259).  Used to hand sympy symbols to the
        user's right-hand-side function as ``p.alpha`` / ``y.x.y.z``."""
        if item_map is None:
            item_map = lambda x: x  # noqa: E731
        pools = {True: np.asarray(flat_subset, dtype=object).ravel(),
                 False: np.asarray(flat_remainder, dtype=object).ravel()}
        cursor = {True: 0, False: 0}
        leaf_by_path = {leaf.path: leaf for leaf in self._leaves}

        def make(name: str, dtype: np.dtype, prefix: Path) -> Any:
            names: List[str] = []
            values: List[Any] = []
            for field, (sub, *_rest) in dtype.fields.items():
                if sub.fields is None:
                    leaf = leaf_by_path[prefix + (field,)]
                    pool = pools[leaf.in_subset]
                    start = cursor[leaf.in_subset]
                    if start + leaf.size > len(pool):
                        raise ValueError("Not enough values for %s" % (leaf.path,))
                    chunk = pool[start:start + leaf.size].reshape(leaf.shape)
                    cursor[leaf.in_subset] = start + leaf.size
                    values.append(item_map(chunk))
                else:
                    values.append(make(field, sub, prefix + (field,)))
                names.append(field)
            cls = dataclasses.make_dataclass(name, names)
            return cls(*values)

        if self.dtype.fields is None:
            result = dataclasses.make_dataclass(dataclass_name, [])()
        else:
            result = make(dataclass_name, self.dtype, ())
        if cursor[True] != len(pools[True]) or cursor[False] != len(pools[False]):
            raise ValueError("Unused values in as_dataclass.")
        return result

    def from_dict(self, vals: Dict[str, Any], out: Optional[np.ndarray] = None) -> np.ndarray:
        if out is None:
            out = np.zeros((1,), dtype=self.dtype)[0]
        _from_dict(out, vals)
        return out

    def subset_from_dict(self, vals: Dict[str, Any], out: Optional[np.ndarray] = None) -> np.ndarray:
        if out is None:
            out = np.zeros((1,), dtype=self.subset_dtype)[0]
        _from_dict(out, vals)
        return out

    def as_dict(self, vals: np.ndarray) -> Dict[str, Any]:
        if vals.dtype != self.dtype:
            raise ValueError('Invalid dtype.')
        return _as_dict(vals)

    def subset_as_dict(self, vals: np.ndarray) -> Dict[str, Any]:
        if vals.dtype != self.subset_dtype:
            raise ValueError('Invalid dtype.')
        return _as_dict(vals)

    @property
    def remainder(self) -> 'DTypeSubset':
        """The complementary subset over the same full layout (dtypesubset.py:283-288)."""
        if self._remainder is None:
            rest = [p for p in self.paths if p not in set(self.subset_paths)]
            # ``self.dims`` already carries (dtype, dim-names) leaves, so no fixed dtype here
            self._remainder = DTypeSubset(self.dims, rest, coords=self.coords)
        return self._remainder
