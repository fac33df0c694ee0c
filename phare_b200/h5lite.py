"""A minimal stand-in for the part of h5py that PHARE's diagnostics use, for machines without HDF5 (this image has neither
h5py nor HighFive).  The reference writes its diagnostics with HighFive (src/diagnostic/detail/h5writer.hpp) and pyphare reads
them with h5py (pyphare/pharesee/hierarchy/fromh5.py): `File(path, mode)`, groups addressed by '/'-separated paths, `.attrs`,
`.keys() / .items() / .values()`, `.name`, datasets that behave like arrays.  This module offers exactly that surface over a
container that is NOT HDF5: a magic line followed by pickled records {path: (attrs, array-or-None)}; a file opened in append mode
adds one record holding only the nodes it touched (a dump costs what it writes, whatever the file already holds), a reader merges
the records in order.  Other ranks put `<file>.rank<r>` pieces next to the file, which a reader merges too (the reference writes one
parallel-HDF5 file from all ranks).

`phare_b200.simulator` writes through real h5py when it is importable and through this module otherwise; registering it as
`sys.modules["h5py"]` (`h5lite.install()`) lets pyphare's own readers (`hierarchy_from(h5_filename=...)`, `Run`) open those files.
"""
import glob
import os
import io
import pickle
import sys

import numpy as np

MAGIC = b"PHB-H5LITE-1\n"
_READ_CACHE = {}  # filename -> ((piece, mtime, size), ...), parsed tree


def _norm(path):
    return "/" + "/".join(k for k in str(path).split("/") if k)


class AttributeManager(dict):
    """h5py's .attrs: values come back as numpy scalars / arrays"""

    def __setitem__(self, key, value):
        if not isinstance(value, (str, bytes)):
            value = np.asarray(value)
            if value.ndim == 0:
                value = value[()]
        super().__setitem__(key, value)

    def create(self, key, data, **kw):
        self[key] = data


class _Node:
    def __init__(self, file, name):
        self.file, self.name = file, name

    @property
    def attrs(self):
        if self.file.mode != "r":
            self.file._dirty.add(self.name)
        return self.file._tree[self.name][0]

    @property
    def parent(self):
        return Group(self.file, _norm(self.name.rsplit("/", 1)[0]))


class Dataset(_Node):
    @property
    def _a(self):
        return self.file._tree[self.name][1]

    shape = property(lambda self: self._a.shape)
    dtype = property(lambda self: self._a.dtype)
    size = property(lambda self: self._a.size)
    ndim = property(lambda self: self._a.ndim)

    def __len__(self):
        return len(self._a)

    def __getitem__(self, idx):
        return self._a[idx]

    def __setitem__(self, idx, value):
        self.file._check_writable()
        self.file._dirty.add(self.name)
        self._a[idx] = value

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._a, dtype=dtype)

    def __iter__(self):
        return iter(self._a)

    def __repr__(self):
        return f'<h5lite dataset "{self.name}": shape {self.shape}, type {self.dtype}>'


class Group(_Node):
    def _abs(self, path):
        path = str(path)
        return _norm(path) if path.startswith("/") else _norm(self.name + "/" + path)

    def _children(self):
        return sorted(self.file._index().get(self.name, ()))

    def _wrap(self, path):
        node = self.file._tree[path]
        return Group(self.file, path) if node[1] is None else Dataset(self.file, path)

    def __contains__(self, path):
        return self._abs(path) in self.file._tree

    def __getitem__(self, path):
        p = self._abs(path)
        if p not in self.file._tree:
            raise KeyError(f"Unable to open object (object '{path}' doesn't exist)")
        return self._wrap(p)

    def get(self, path, default=None):
        return self[path] if path in self else default

    def keys(self):
        return self._children()

    def __iter__(self):
        return iter(self._children())

    def __len__(self):
        return len(self._children())

    def values(self):
        return [self[k] for k in self._children()]

    def items(self):
        return [(k, self[k]) for k in self._children()]

    def _make_parents(self, p):
        parts = [k for k in p.split("/") if k]
        for i in range(len(parts)):
            k = "/" + "/".join(parts[:i])
            if k not in self.file._tree:
                self.file._tree[k] = (AttributeManager(), None)
                self.file._dirty.add(k)

    def create_group(self, path):
        self.file._check_writable()
        p = self._abs(path)
        if p in self.file._tree:
            raise ValueError(f"Unable to create group (name already exists): {p}")
        self._make_parents(p)
        self.file._tree[p] = (AttributeManager(), None)
        self.file._dirty.add(p)
        return Group(self.file, p)

    def require_group(self, path):
        return self[path] if path in self else self.create_group(path)

    def create_dataset(self, path, shape=None, dtype=None, data=None, **kw):
        self.file._check_writable()
        p = self._abs(path)
        if p in self.file._tree:
            raise ValueError(f"Unable to create dataset (name already exists): {p}")
        a = np.zeros(shape, dtype or np.float64) if data is None else np.array(data, dtype=dtype)
        self._make_parents(p)
        self.file._tree[p] = (AttributeManager(), a)
        self.file._dirty.add(p)
        return Dataset(self.file, p)

    def __setitem__(self, path, data):
        self.create_dataset(path, data=data)

    def __delitem__(self, path):
        self.file._check_writable()
        p = self._abs(path)
        for k in [k for k in self.file._tree if k == p or k.startswith(p + "/")]:
            del self.file._tree[k]
        self.file._rewrite, self.file._kids = True, None

    def visit(self, fn):
        prefix = self.name.rstrip("/") + "/"
        for k in sorted(self.file._tree):
            if k.startswith(prefix) and k != prefix:
                r = fn(k[len(prefix):])
                if r is not None:
                    return r

    def __repr__(self):
        return f'<h5lite group "{self.name}" ({len(self)} members)>'


class File(Group):
    """h5py.File(filename, mode): 'r' (default), 'r+', 'a', 'w', 'w-' / 'x'"""

    def __init__(self, filename, mode="r", **kw):
        self.filename, self.mode = str(filename), mode
        self._tree, self._open = {"/": (AttributeManager(), None)}, True
        self._dirty, self._rewrite = {"/"}, False
        self._kids, self._kids_n = None, -1
        exists = os.path.exists(self.filename)
        if mode in ("r", "r+") and not exists:
            raise FileNotFoundError(f"Unable to open file (unable to open file: name = '{filename}')")
        if mode in ("w-", "x") and exists:
            raise FileExistsError(f"Unable to create file (file exists): {filename}")
        if mode in ("r", "r+", "a") and exists:
            pieces = [self.filename] + (sorted(glob.glob(glob.escape(self.filename) + ".rank*")) if mode == "r" else [])
            sig = tuple((q, os.stat(q).st_mtime_ns, os.stat(q).st_size) for q in pieces)
            cached = _READ_CACHE.get(self.filename) if mode == "r" else None
            if cached is not None and cached[0] == sig:
                self._tree = cached[1]  # read-only handles of an unchanged file share the parsed tree
            else:
                for piece in pieces:
                    self._merge(piece)
                if mode == "r":
                    if len(_READ_CACHE) >= 64:
                        _READ_CACHE.pop(next(iter(_READ_CACHE)))
                    _READ_CACHE[self.filename] = (sig, self._tree)
            self._dirty = set()
        else:
            self._rewrite = True  # a new file: magic + first record
        super().__init__(self, "/")

    def _index(self):
        """parent path -> names of its children, rebuilt when the tree changed size"""
        if self._kids is None or self._kids_n != len(self._tree):
            kids = {}
            for k in self._tree:
                if k != "/":
                    parent, _, name = k.rpartition("/")
                    kids.setdefault(parent or "/", set()).add(name)
            self._kids, self._kids_n = kids, len(self._tree)
        return self._kids

    def _merge(self, piece):
        with open(piece, "rb") as f:
            if f.read(len(MAGIC)) != MAGIC:
                raise OSError(f"Unable to open file (not an h5lite container): {piece}")
            while True:
                try:
                    record = _SafeUnpickler(f).load()
                except EOFError:
                    break
                for k, (attrs, a) in record.items():
                    if k in self._tree and a is None and self._tree[k][1] is None:
                        self._tree[k][0].update(attrs)
                    else:
                        self._tree[k] = (AttributeManager(attrs), a)

    def _check_writable(self):
        if not self._open or self.mode == "r":
            raise OSError("h5lite file is not open for writing")

    def flush(self):
        if self.mode == "r" or not self._open:
            return
        if self._rewrite:
            tmp = self.filename + ".tmp"
            with open(tmp, "wb") as f:
                f.write(MAGIC)
                pickle.dump({k: (dict(v[0]), v[1]) for k, v in self._tree.items()}, f, protocol=pickle.HIGHEST_PROTOCOL)
            os.replace(tmp, self.filename)
        elif self._dirty:
            with open(self.filename, "ab") as f:
                pickle.dump({k: (dict(self._tree[k][0]), self._tree[k][1]) for k in self._dirty if k in self._tree}, f,
                            protocol=pickle.HIGHEST_PROTOCOL)
        self._dirty, self._rewrite = set(), False

    def release_datasets(self):
        """writer-side helper for a file kept open across many dumps: flush, then forget the arrays already on disk (the
        groups stay, so later dumps keep appending under them); the file must not be read through this handle afterwards"""
        self.flush()
        for k in [k for k, v in self._tree.items() if v[1] is not None]:
            del self._tree[k]
        self._kids = None

    def close(self):
        self.flush()
        self._open = False

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __bool__(self):
        return self._open

    def __repr__(self):
        return f'<h5lite file "{os.path.basename(self.filename)}" (mode {self.mode})>'


class _SafeUnpickler(pickle.Unpickler):
    """records hold only dicts, strings, numbers and numpy arrays / scalars: anything else in a container (a file from
    another user or machine may have been tampered with) is refused instead of being imported and called"""
    ALLOWED = {("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
               ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"),
               ("numpy", "ndarray"), ("numpy", "dtype"), ("numpy.core.numeric", "_frombuffer"),
               ("numpy._core.numeric", "_frombuffer"), ("builtins", "dict"), ("builtins", "list"), ("builtins", "tuple"),
               ("builtins", "set"), ("builtins", "frozenset"), ("builtins", "bytes"), ("builtins", "bytearray"),
               ("builtins", "complex"), ("builtins", "slice")}

    def find_class(self, module, name):
        if (module, name) in self.ALLOWED or (module.startswith("numpy.dtypes") and name.endswith("DType")):
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"h5lite: refusing to load {module}.{name} from a container")


def is_h5lite(filename):
    with open(filename, "rb") as f:
        return f.read(len(MAGIC)) == MAGIC


def install():
    """make `import h5py` resolve to this module (only when the real h5py is absent)"""
    try:
        import h5py  # noqa: F401
        if getattr(h5py, "__name__", "") != __name__ and hasattr(h5py, "File") and not hasattr(h5py, "_mock_name"):
            return h5py
    except ImportError:
        pass
    sys.modules["h5py"] = sys.modules[__name__]
    return sys.modules[__name__]
