"""CLAM-style HDF5 bag files without h5py: a thin wrapper over the native reader in libmoc_b200.so.

The reference opens ``h5_files/<slide>.h5`` with h5py and reads ``['features']`` / ``['coords']`` in full
(datasets/dataset_generic.py:424-430); neither h5py nor libhdf5 exists in this image, so the subset of the HDF5
format those files use is parsed natively (moc_b200/csrc/h5_reader.cu).  ``H5File`` mimics the slice of h5py the
reference touches: ``with H5File(path) as f: f['features'][:]``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check

_INT = {1: np.int8, 2: np.int16, 4: np.int32, 8: np.int64}
_FLT = {2: np.float16, 4: np.float32, 8: np.float64}


class _Dataset:
    def __init__(self, f: "H5File", name: str):
        self._f, self.name = f, name
        rank, dims, cls, es = C.c_int(), (C.c_int64 * 4)(), C.c_int(), C.c_int()
        check(_lib.load().moc_h5_dataset_info(f._h, name.encode(), C.byref(rank), dims, C.byref(cls), C.byref(es)))
        self.shape = tuple(int(dims[i]) for i in range(rank.value))
        table = _FLT if cls.value == 1 else _INT
        if es.value not in table:
            raise _lib.MocError(_lib.E_SHAPE, "dataset %s has an unsupported element size %d" % (name, es.value))
        self.dtype = np.dtype(table[es.value])

    @property
    def nbytes(self) -> int:
        return int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize

    def read_into(self, ptr: int, nbytes: int) -> None:
        """Dense row-major copy of the dataset into host memory at ``ptr`` (e.g. a pinned staging buffer)."""
        check(_lib.load().moc_h5_read(self._f._h, self.name.encode(), ptr, nbytes))

    def __getitem__(self, key):
        out = np.empty(self.shape, dtype=self.dtype)
        self.read_into(out.ctypes.data, out.nbytes)
        return out[key]


class H5File:
    def __init__(self, path: str, mode: str = "r"):
        if mode != "r":
            raise ValueError("moc_b200.h5bag.H5File is read-only")
        self._h = C.c_void_p()
        check(_lib.load().moc_h5_open(str(path).encode(), C.byref(self._h)))

    def __getitem__(self, name: str) -> _Dataset:
        return _Dataset(self, name)

    def close(self) -> None:
        if self._h:
            _lib.load().moc_h5_close(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
