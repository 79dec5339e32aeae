"""ctypes binding of oracle/_ref/libxnref_model.so -- the reference's own src/model code
(Grid::load_tiff via the real libtiff, build_octree, Octree::save_svo/load_svo).
TEST INFRASTRUCTURE ONLY; present only where /root/reference existed at build time."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libxnref_model.so")


class BuildStats(C.Structure):
    _fields_ = [("total_leaves", C.c_uint64), ("unique_leaves", C.c_uint64),
                ("total_nodes", C.c_uint64), ("depth", C.c_uint64)]


def available() -> bool:
    return os.path.exists(SO)


_lib = None


def lib():
    global _lib
    if _lib is None:
        # Pillow's bundled libtiff (and its zstd/lzma/jpeg deps) resolve once PIL's C core is loaded
        from PIL import _imaging  # noqa: F401

        _lib = C.CDLL(SO)
    return _lib


def _stats(st):
    return {k: getattr(st, k) for k, _ in BuildStats._fields_}


def convert_mem(grid, svo_path, *, chan_diff=None, std_dev=None, type=0):
    grid = np.ascontiguousarray(grid, dtype=np.uint8)
    nz, ny, nx, _ = grid.shape
    heur, param = (1, float(std_dev)) if std_dev is not None else (0, float(chan_diff or 0))
    st = BuildStats()
    f = lib().xnref_convert_mem
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_char_p, C.c_int, C.c_double, C.c_int,
                  C.POINTER(BuildStats)]
    rc = f(grid.ctypes.data, nx, ny, nz, os.fsencode(svo_path), heur, param, type, C.byref(st))
    if rc != 0:
        raise RuntimeError("reference convert failed")
    return _stats(st)


def convert(tif_path, svo_path, *, chan_diff=None, std_dev=None, type=0):
    heur, param = (1, float(std_dev)) if std_dev is not None else (0, float(chan_diff or 0))
    st = BuildStats()
    f = lib().xnref_convert
    f.restype = C.c_int
    f.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_double, C.c_int, C.POINTER(BuildStats)]
    rc = f(os.fsencode(tif_path), os.fsencode(svo_path), heur, param, type, C.byref(st))
    if rc != 0:
        raise RuntimeError("reference convert failed")
    return _stats(st)


def load_tiff(tif_path):
    """Grid::load_tiff -> uint8 array (nz, ny, nx, 4)."""
    f = lib().xnref_load_tiff
    f.restype = C.c_int
    f.argtypes = [C.c_char_p, C.POINTER(C.c_uint64 * 3), C.c_void_p, C.c_uint64]
    dims = (C.c_uint64 * 3)()
    if f(os.fsencode(tif_path), C.byref(dims), None, 0) != 0:
        raise RuntimeError("reference load_tiff failed")
    nx, ny, nz = dims
    out = np.empty((nz, ny, nx, 4), dtype=np.uint8)
    if f(os.fsencode(tif_path), C.byref(dims), out.ctypes.data, out.nbytes) != 0:
        raise RuntimeError("reference load_tiff failed")
    return out


def svo_roundtrip(src, dst):
    f = lib().xnref_svo_roundtrip
    f.restype = C.c_int
    f.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    side, count = C.c_uint64(), C.c_uint64()
    rc = f(os.fsencode(src), os.fsencode(dst) if dst else None, C.byref(side), C.byref(count))
    if rc != 0:
        raise RuntimeError("reference load_svo rejected the file")
    return side.value, count.value
