"""ctypes binding of oracle/_ref/libxnref_glsl.so -- the reference's own shader text
compiled as C++ (oracle/glsl_shim).  TEST INFRASTRUCTURE ONLY; exists only where
/root/reference was present at build time (this container), so tests that need it
skip when the library is absent and rely on the committed tests/golden fixtures."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libxnref_glsl.so")

_NAMES = {"dda": "dda", "svo-naive": "svo_naive", "svo-df": "svo_df", "esvo": "esvo", "svo-rope": "svo_rope"}


class Args(C.Structure):
    _fields_ = [
        ("forward", C.c_float * 3), ("up", C.c_float * 3), ("translation", C.c_float * 3),
        ("out_ox", C.c_int32), ("out_oy", C.c_int32), ("out_w", C.c_uint32), ("out_h", C.c_uint32),
        ("disp_ox", C.c_int32), ("disp_oy", C.c_int32), ("disp_w", C.c_uint32), ("disp_h", C.c_uint32),
        ("voxel_ratio", C.c_float * 3), ("model_dim", C.c_uint32 * 3), ("emission_coeff", C.c_float),
        ("volume", C.c_void_p), ("nx", C.c_uint64), ("ny", C.c_uint64), ("nz", C.c_uint64),
        ("rgba_out", C.c_void_p),
    ]


def available() -> bool:
    return os.path.exists(SO)


_lib = None


def render(traversal: str, *, grid=None, nodes=None, side=None, camera, output, display=None,
           voxel_ratio=(1, 1, 1), emission=1.0):
    """Same calling convention as oracle.xo.render; returns rgba uint8 (h, w, 4)."""
    global _lib
    if _lib is None:
        _lib = C.CDLL(SO)
    display = display or output
    a = Args()
    ratio = np.asarray(voxel_ratio, dtype=np.float32)
    a.forward[:] = [float(x) for x in camera[0]]
    a.up[:] = [float(x) for x in camera[1]]
    # src/render/Renderer.cpp:62 -- the host pre-divides the translation (binary32)
    tr = np.asarray(camera[2], dtype=np.float32) / ratio
    a.translation[:] = [float(x) for x in tr]
    a.out_ox, a.out_oy, a.out_w, a.out_h = output
    a.disp_ox, a.disp_oy, a.disp_w, a.disp_h = display
    a.voxel_ratio[:] = [float(x) for x in ratio]
    a.emission_coeff = float(emission)
    if traversal == "dda":
        grid = np.ascontiguousarray(grid, dtype=np.uint8)
        nz, ny, nx, _ = grid.shape
        a.volume = grid.ctypes.data
        a.nx, a.ny, a.nz = nx, ny, nz
        a.model_dim[:] = [nx, ny, nz]
    else:
        nodes = np.ascontiguousarray(nodes)
        assert nodes.dtype.itemsize == 40
        a.volume = nodes.ctypes.data
        a.model_dim[:] = [side, side, side]
    h, w = output[3], output[2]
    rgba = np.zeros((h, w), dtype=np.uint32)
    a.rgba_out = rgba.ctypes.data
    fn = getattr(_lib, "xnref_render_" + _NAMES[traversal])
    fn.restype = C.c_int
    fn.argtypes = [C.POINTER(Args)]
    if fn(C.byref(a)) != 0:
        raise RuntimeError("xnref_render failed")
    return rgba.view(np.uint8).reshape(h, w, 4)
