"""ctypes binding of the CPU oracle (oracle/libxn_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Never by xenodon_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

DDA, SVO_NAIVE, ESVO, SVO_DF, SVO_ROPE = range(5)
TRAVERSALS = {"dda": DDA, "svo-naive": SVO_NAIVE, "esvo": ESVO, "svo-df": SVO_DF, "svo-rope": SVO_ROPE}
TYPE_SPARSE, TYPE_DAG, TYPE_ROPE = range(3)
HEUR_CHAN_DIFF, HEUR_STD_DEV = range(2)

NODE_DTYPE = np.dtype([("children", "<u4", (8,)), ("color", "<u4"), ("is_leaf_depth", "<u4")])
assert NODE_DTYPE.itemsize == 40


class Rect(C.Structure):
    _fields_ = [("ox", C.c_int32), ("oy", C.c_int32), ("w", C.c_uint32), ("h", C.c_uint32)]


class Params(C.Structure):
    _fields_ = [("voxel_ratio", C.c_float * 3), ("model_dim", C.c_uint32 * 3), ("emission_coeff", C.c_float)]


class Camera(C.Structure):
    _fields_ = [("forward", C.c_float * 3), ("up", C.c_float * 3), ("translation", C.c_float * 3)]


class Volume(C.Structure):
    _fields_ = [
        ("grid", C.c_void_p),
        ("nx", C.c_uint64),
        ("ny", C.c_uint64),
        ("nz", C.c_uint64),
        ("nodes", C.c_void_p),
        ("num_nodes", C.c_uint64),
    ]


class BuildStats(C.Structure):
    _fields_ = [
        ("total_leaves", C.c_uint64),
        ("unique_leaves", C.c_uint64),
        ("total_nodes", C.c_uint64),
        ("depth", C.c_uint64),
    ]


def build(force: bool = False) -> str:
    """Compile oracle/libxn_oracle.so if missing (gcc, see oracle/Makefile)."""
    so = os.path.join(_HERE, "libxn_oracle.so")
    src = os.path.join(_HERE, "xn_oracle.c")
    if force or not os.path.exists(so) or (
        os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)
    ):
        subprocess.run(["make", "-C", _HERE, "libxn_oracle.so"], check=True, capture_output=True)
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.xo_render.restype = C.c_int
        _lib.xo_render.argtypes = [
            C.c_int, C.POINTER(Volume), C.POINTER(Params), C.POINTER(Camera), C.POINTER(Rect), C.POINTER(Rect),
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
        ]
        _lib.xo_build_octree.restype = C.c_int
        _lib.xo_build_octree.argtypes = [
            C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_double, C.c_int,
            C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(BuildStats),
        ]
        _lib.xo_generate_ropes.restype = None
        _lib.xo_generate_ropes.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
        _lib.xo_free.restype = None
        _lib.xo_free.argtypes = [C.c_void_p]
    return _lib


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def render(traversal, *, grid=None, nodes=None, side=None, camera, output, display=None,
           voxel_ratio=(1, 1, 1), emission=1.0, threads=0, want_stats=True):
    """Render one region.  grid: uint8 array (nz, ny, nx, 4); nodes: NODE_DTYPE array.

    camera = (forward, up, translation) in user units; output/display = (ox, oy, w, h).
    Returns (rgba uint8 (h, w, 4), steps uint32 (h, w) | None, bytes uint64 (h, w) | None).
    """
    if isinstance(traversal, str):
        traversal = TRAVERSALS[traversal]
    display = display or output
    vol = Volume()
    if traversal == DDA:
        grid = np.ascontiguousarray(grid, dtype=np.uint8)
        nz, ny, nx, _ = grid.shape
        vol.grid = grid.ctypes.data
        vol.nx, vol.ny, vol.nz = nx, ny, nz
        model_dim = (nx, ny, nz)
    else:
        nodes = np.ascontiguousarray(nodes, dtype=NODE_DTYPE)
        vol.nodes = nodes.ctypes.data
        vol.num_nodes = len(nodes)
        model_dim = (side, side, side)
    p = Params(_f3(voxel_ratio), (C.c_uint32 * 3)(*model_dim), float(emission))
    cam = Camera(_f3(camera[0]), _f3(camera[1]), _f3(camera[2]))
    out = Rect(*output)
    disp = Rect(*display)
    h, w = output[3], output[2]
    rgba = np.empty((h, w), dtype=np.uint32)
    steps = np.empty((h, w), dtype=np.uint32) if want_stats else None
    nbytes = np.empty((h, w), dtype=np.uint64) if want_stats else None
    rc = lib().xo_render(
        traversal, C.byref(vol), C.byref(p), C.byref(cam), C.byref(out), C.byref(disp),
        rgba.ctypes.data, steps.ctypes.data if want_stats else None,
        nbytes.ctypes.data if want_stats else None, threads,
    )
    if rc != 0:
        raise ValueError("xo_render: bad arguments")
    return rgba.view(np.uint8).reshape(h, w, 4), steps, nbytes


def build_octree(grid, *, chan_diff=None, std_dev=None, type=TYPE_SPARSE):
    """`xenodon convert` restated.  Returns (nodes, side, stats dict)."""
    grid = np.ascontiguousarray(grid, dtype=np.uint8)
    nz, ny, nx, _ = grid.shape
    if std_dev is not None:
        heur, param = HEUR_STD_DEV, float(std_dev)
    else:
        heur, param = HEUR_CHAN_DIFF, float(chan_diff or 0)
    out = C.c_void_p()
    count = C.c_uint64()
    side = C.c_uint64()
    st = BuildStats()
    rc = lib().xo_build_octree(grid.ctypes.data, nx, ny, nz, heur, param, type,
                               C.byref(out), C.byref(count), C.byref(side), C.byref(st))
    if rc != 0:
        raise ValueError("xo_build_octree: bad arguments")
    buf = (C.c_char * (count.value * 40)).from_address(out.value)
    nodes = np.frombuffer(buf, dtype=NODE_DTYPE).copy()
    lib().xo_free(out)
    stats = {k: getattr(st, k) for k, _ in BuildStats._fields_}
    return nodes, side.value, stats


def generate_ropes(nodes, side):
    nodes = np.ascontiguousarray(nodes, dtype=NODE_DTYPE).copy()
    lib().xo_generate_ropes(nodes.ctypes.data, len(nodes), side)
    return nodes
