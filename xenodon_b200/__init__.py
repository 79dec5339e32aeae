"""xenodon_b200 -- B200-native volume ray traversal (Xenodon's hot path) behind a C ABI.

This package is the Python host-side mirror of the reference's interface for the path
(reference src/render, src/model, src/backend/headless): thin ctypes wrappers over
`libxenodon_b200.so` (include/xenodon_b200.h).  All compute happens in the CUDA library;
there is no CPU fallback, and importing the library without having built it fails loudly.

Names follow the reference: Grid.load_tiff, Octree.load_svo/save_svo, build_octree,
HeadlessConfig, RenderContext-style display rectangle, MultiplexRenderer.render/stats.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
# XN_LIBRARY: an alternative build of the same library (tuning A/B runs, xenodon_b200.build --variant)
LIB_PATH = os.environ.get("XN_LIBRARY") or os.path.join(_PKG, "libxenodon_b200.so")
CLI_PATH = os.path.join(_PKG, "bin", "xenodon")

DDA, SVO_NAIVE, ESVO, SVO_DF, SVO_ROPE = range(5)
TRAVERSALS = {"dda": DDA, "svo-naive": SVO_NAIVE, "esvo": ESVO, "svo-df": SVO_DF, "svo-rope": SVO_ROPE}
TYPE_SPARSE, TYPE_DAG, TYPE_ROPE = range(3)
HEUR_CHAN_DIFF, HEUR_STD_DEV = range(2)
SYNTH_BUNNY, SYNTH_TNG = 0, 1
LAYOUT_AUTO, LAYOUT_LINEAR, LAYOUT_BRICKED, LAYOUT_TEXTURE = range(4)

NODE_DTYPE = np.dtype([("children", "<u4", (8,)), ("color", "<u4"), ("is_leaf_depth", "<u4")])
assert NODE_DTYPE.itemsize == 40


class XenodonError(RuntimeError):
    """Failure reported by the C ABI; .status is the xn_status code."""

    def __init__(self, status: int, message: str):
        super().__init__(message)
        self.status = status


class Rect(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("w", C.c_uint32), ("h", C.c_uint32)]

    def astuple(self):
        return (self.x, self.y, self.w, self.h)


class RenderStats(C.Structure):
    _fields_ = [("total_rays", C.c_uint64), ("outputs", C.c_uint64), ("total_render_time", C.c_double),
                ("max_render_time", C.c_double), ("min_render_time", C.c_double)]

    def mrays_per_s(self) -> float:
        return self.total_rays / (self.total_render_time * 1000.0)


class BuildStats(C.Structure):
    _fields_ = [("total_leaves", C.c_uint64), ("unique_leaves", C.c_uint64), ("total_nodes", C.c_uint64),
                ("depth", C.c_uint64)]


class HeadlessDevice(C.Structure):
    _fields_ = [("vkindex", C.c_uint32), ("region", Rect)]


_lib = None

_PROTOTYPES = {
    "xn_last_error": (C.c_char_p, []),
    "xn_version": (C.c_char_p, []),
    "xn_traversal_from_name": (C.c_int, [C.c_char_p]),
    "xn_traversal_name": (C.c_char_p, [C.c_int]),
    "xn_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "xn_device_name": (C.c_int, [C.c_int, C.c_char_p, C.c_size_t]),
    "xn_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "xn_ctx_destroy": (C.c_int, [C.c_void_p]),
    "xn_ctx_device": (C.c_int, [C.c_void_p]),
    "xn_upload_grid": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64]),
    "xn_upload_svo": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64]),
    "xn_tiff_stream_info": (C.c_int, [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]),
    "xn_upload_grid_tiff": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_uint64), C.POINTER(C.c_double)]),
    "xn_upload_grid_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64]),
    "xn_upload_svo_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64]),
    "xn_convert_resident_grid": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p),
                                           C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(BuildStats)]),
    "xn_convert_resident_grid_ex": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int, C.POINTER(C.c_void_p),
                                              C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(BuildStats)]),
    "xn_synth_grid_device": (C.c_int, [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32]),
    "xn_synth_grid_host": (C.c_int, [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p]),
    "xn_download_grid": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "xn_set_grid_layout": (C.c_int, [C.c_void_p, C.c_int]),
    "xn_grid_layout": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_uint64)]),
    "xn_brick_layout": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_uint64)]),
    "xn_brick_indices": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p]),
    "xn_set_target": (C.c_int, [C.c_void_p, C.POINTER(Rect), C.POINTER(Rect)]),
    "xn_set_params": (C.c_int, [C.c_void_p, C.POINTER(C.c_float * 3), C.POINTER(C.c_uint32 * 3), C.c_float]),
    "xn_set_precision": (C.c_int, [C.c_void_p, C.c_int]),
    "xn_set_interleave": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32]),
    "xn_owned_rays": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "xn_set_target_buffer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "xn_render": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float * 3), C.POINTER(C.c_float * 3),
                            C.POINTER(C.c_float * 3)]),
    "xn_sync": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "xn_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "xn_render_download_async": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float * 3), C.POINTER(C.c_float * 3),
                                           C.POINTER(C.c_float * 3), C.c_void_p]),
    "xn_render_download_to": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float * 3), C.POINTER(C.c_float * 3),
                                        C.POINTER(C.c_float * 3), C.c_void_p, C.c_size_t]),
    "xn_signal_after_copy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32]),
    "xn_host_register": (C.c_int, [C.c_void_p, C.c_size_t]),
    "xn_host_unregister": (C.c_int, [C.c_void_p]),
    "xn_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "xn_host_free": (C.c_int, [C.c_void_p]),
    "xn_mark": (C.c_int, [C.c_void_p, C.c_int]),
    "xn_mark_elapsed": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "xn_launch_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "xn_render_stats_pass": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float * 3), C.POINTER(C.c_float * 3),
                                       C.POINTER(C.c_float * 3), C.c_void_p, C.c_void_p,
                                       C.POINTER(C.c_uint64 * 2)]),
    "xn_render_touch_pass": (C.c_int, [C.c_void_p, C.POINTER(C.c_float * 3), C.POINTER(C.c_float * 3),
                                       C.POINTER(C.c_float * 3), C.c_int, C.POINTER(C.c_uint64 * 2)]),
    "xn_frame_gather": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.POINTER(Rect)]),
    "xn_frame_buffer_create": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p), C.c_void_p]),
    "xn_frame_buffer_open": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "xn_frame_buffer_close": (C.c_int, [C.c_void_p, C.c_void_p]),
    "xn_frame_buffer_read": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]),
    "xn_frame_buffer_read_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]),
    "xn_copy_sync": (C.c_int, [C.c_void_p]),
    "xn_tiff_info": (C.c_int, [C.c_char_p, C.POINTER(C.c_uint64 * 3)]),
    "xn_tiff_read": (C.c_int, [C.c_char_p, C.c_void_p, C.c_uint64]),
    "xn_tiff_write": (C.c_int, [C.c_char_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int]),
    "xn_svo_info": (C.c_int, [C.c_char_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "xn_svo_read": (C.c_int, [C.c_char_p, C.c_void_p, C.c_uint64]),
    "xn_svo_write": (C.c_int, [C.c_char_p, C.c_void_p, C.c_uint64, C.c_uint64]),
    "xn_build_octree": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_double, C.c_int,
                                  C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                  C.POINTER(BuildStats)]),
    "xn_free": (None, [C.c_void_p]),
    "xn_headless_config_parse": (C.c_int, [C.c_char_p, C.POINTER(HeadlessDevice), C.c_int, C.POINTER(C.c_int)]),
    "xn_camera_script_parse": (C.c_int, [C.c_char_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "xn_stats_write": (C.c_int, [C.c_char_p, C.POINTER(RenderStats), C.c_uint64, C.c_double]),
    "xn_png_write": (C.c_int, [C.c_char_p, C.c_void_p, C.c_uint32, C.c_uint32]),
}


def lib():
    """The loaded C ABI.  Raises if the extension has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build the CUDA extension first "
                "(python -m xenodon_b200.build, or __graft_entry__.build()); there is no CPU fallback"
            )
        handle = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in _PROTOTYPES.items():
            fn = getattr(handle, name)  # AttributeError if the library does not export it
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def _check(rc: int):
    if rc != 0:
        raise XenodonError(rc, lib().xn_last_error().decode("utf-8", "replace"))


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def device_count() -> int:
    n = C.c_int()
    _check(lib().xn_device_count(C.byref(n)))
    return n.value


def device_name(i: int) -> str:
    buf = C.create_string_buffer(256)
    _check(lib().xn_device_name(i, buf, 256))
    return buf.value.decode()


# ----------------------------------------------------------------------------------------
# model: Grid / Octree (reference src/model)
# ----------------------------------------------------------------------------------------
class Grid:
    """RGBA8 voxel grid, `data[z, y, x] = (r, g, b, a)` (x fastest, reference Grid.h:50-52)."""

    def __init__(self, data: np.ndarray):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        if data.ndim != 4 or data.shape[3] != 4:
            raise ValueError("grid must have shape (nz, ny, nx, 4)")
        self.data = data

    @property
    def dimensions(self):
        nz, ny, nx, _ = self.data.shape
        return (nx, ny, nz)

    @staticmethod
    def load_tiff(path) -> "Grid":
        dims = (C.c_uint64 * 3)()
        _check(lib().xn_tiff_info(os.fsencode(path), C.byref(dims)))
        nx, ny, nz = dims
        out = np.empty((nz, ny, nx, 4), dtype=np.uint8)
        _check(lib().xn_tiff_read(os.fsencode(path), out.ctypes.data, out.nbytes))
        return Grid(out)

    def save_tiff(self, path, bigtiff: bool = True):
        nx, ny, nz = self.dimensions
        _check(lib().xn_tiff_write(os.fsencode(path), self.data.ctypes.data, nx, ny, nz, int(bigtiff)))

    @staticmethod
    def synthetic(kind: int, nx: int, ny: int, nz: int, seed: int = 1729) -> "Grid":
        out = np.empty((nz, ny, nx, 4), dtype=np.uint8)
        _check(lib().xn_synth_grid_host(kind, nx, ny, nz, seed, out.ctypes.data))
        return Grid(out)


class Octree:
    def __init__(self, nodes: np.ndarray, side: int):
        self.nodes = np.ascontiguousarray(nodes, dtype=NODE_DTYPE)
        self.side = int(side)

    @staticmethod
    def load_svo(path) -> "Octree":
        side, count = C.c_uint64(), C.c_uint64()
        _check(lib().xn_svo_info(os.fsencode(path), C.byref(side), C.byref(count)))
        nodes = np.empty(count.value, dtype=NODE_DTYPE)
        _check(lib().xn_svo_read(os.fsencode(path), nodes.ctypes.data, count.value))
        return Octree(nodes, side.value)

    def save_svo(self, path):
        _check(lib().xn_svo_write(os.fsencode(path), self.nodes.ctypes.data, len(self.nodes), self.side))


def tiff_stream_info(path) -> dict:
    """How the ingest pipeline will take a TIFF (csrc/host/xn_tiff.cpp tiff_plan)."""
    ok, fmt, runs = C.c_int(), (C.c_uint32 * 5)(), C.c_uint64()
    _check(lib().xn_tiff_stream_info(os.fsencode(path), C.byref(ok), fmt, C.byref(runs)))
    return {"streamable": bool(ok.value), "samples": fmt[0], "photometric": fmt[1], "has_alpha": bool(fmt[2]),
            "unassociated": bool(fmt[3]), "flip": bool(fmt[4] & 1), "mirror": bool(fmt[4] & 2),
            "runs": runs.value}


def brick_layout(nx: int, ny: int, nz: int, top: int = -1) -> dict:
    """Bricked residency layout of an nx*ny*nz grid (csrc/xn_brick.h)."""
    d = (C.c_uint64 * 8)()
    _check(lib().xn_brick_layout(nx, ny, nz, top, d))
    return {"mask": tuple(d[0:3]), "hs": tuple(d[3:6]), "top": int(d[6]), "total": int(d[7])}


def brick_indices(nx: int, ny: int, nz: int, xyz, top: int = -1) -> np.ndarray:
    xyz = np.ascontiguousarray(xyz, dtype=np.int32).reshape(-1, 3)
    out = np.empty(len(xyz), dtype=np.uint64)
    _check(lib().xn_brick_indices(nx, ny, nz, top, xyz.ctypes.data, len(xyz), out.ctypes.data))
    return out


def build_octree(grid: Grid, *, chan_diff=None, std_dev=None, type: int = TYPE_SPARSE):
    """`xenodon convert` (reference OctreeConstruction.h:226-237).  Returns (Octree, stats dict)."""
    if chan_diff is not None and std_dev is not None:
        raise ValueError("--std-dev and --chan-diff are mutually exclusive")
    heur, param = (HEUR_STD_DEV, float(std_dev)) if std_dev is not None else (HEUR_CHAN_DIFF, float(chan_diff or 0))
    nx, ny, nz = grid.dimensions
    out, count, side, st = C.c_void_p(), C.c_uint64(), C.c_uint64(), BuildStats()
    _check(lib().xn_build_octree(grid.data.ctypes.data, nx, ny, nz, heur, param, type, C.byref(out),
                                 C.byref(count), C.byref(side), C.byref(st)))
    try:
        buf = (C.c_char * (count.value * 40)).from_address(out.value)
        nodes = np.frombuffer(buf, dtype=NODE_DTYPE).copy()
    finally:
        lib().xn_free(out)
    return Octree(nodes, side.value), {k: getattr(st, k) for k, _ in BuildStats._fields_}


# ----------------------------------------------------------------------------------------
# text formats
# ----------------------------------------------------------------------------------------
def parse_headless_config(text: str):
    """-> list of (vkindex, (x, y, w, h)) (reference HeadlessConfig.cpp:5-28)."""
    n = C.c_int()
    _check(lib().xn_headless_config_parse(text.encode(), None, 0, C.byref(n)))
    arr = (HeadlessDevice * n.value)()
    _check(lib().xn_headless_config_parse(text.encode(), arr, n.value, C.byref(n)))
    return [(d.vkindex, d.region.astuple()) for d in arr]


def parse_camera_script(text: str) -> np.ndarray:
    """-> float32 array (frames, 3, 3): forward, up, translation per frame."""
    n = C.c_int()
    _check(lib().xn_camera_script_parse(text.encode(), None, 0, C.byref(n)))
    out = np.empty((n.value, 3, 3), dtype=np.float32)
    _check(lib().xn_camera_script_parse(text.encode(), out.ctypes.data, n.value, C.byref(n)))
    return out


def write_stats(path, frames, wall_seconds: float):
    arr = (RenderStats * len(frames))(*frames)
    _check(lib().xn_stats_write(os.fsencode(path), arr, len(frames), float(wall_seconds)))


def write_png(path, rgba: np.ndarray):
    rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
    h, w, _ = rgba.shape
    _check(lib().xn_png_write(os.fsencode(path), rgba.ctypes.data, w, h))


def rect_union(rects):
    x0 = min(r[0] for r in rects)
    y0 = min(r[1] for r in rects)
    x1 = max(r[0] + r[2] for r in rects)
    y1 = max(r[1] + r[3] for r in rects)
    return (x0, y0, x1 - x0, y1 - y0)


# ----------------------------------------------------------------------------------------
# render: one Context per (device, region) -- the reference's Renderer + HeadlessOutput
# ----------------------------------------------------------------------------------------
class Context:
    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        _check(lib().xn_ctx_create(device, C.byref(self._h)))
        self.device = device
        self.output = None
        self.display = None

    def close(self):
        if self._h:
            lib().xn_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload_grid(self, grid: Grid):
        nx, ny, nz = grid.dimensions
        _check(lib().xn_upload_grid(self._h, grid.data.ctypes.data, nx, ny, nz))
        self.model_dim = (nx, ny, nz)

    def upload_grid_tiff(self, path: str) -> float:
        """Pipelined TIFF ingest (pinned staging, decode on the device); returns the wall seconds."""
        dims, secs = (C.c_uint64 * 3)(), C.c_double()
        _check(lib().xn_upload_grid_tiff(self._h, os.fsencode(path), dims, C.byref(secs)))
        self.model_dim = tuple(int(d) for d in dims)
        return secs.value

    def upload_grid_device(self, device_ptr: int, nx: int, ny: int, nz: int):
        _check(lib().xn_upload_grid_device(self._h, device_ptr, nx, ny, nz))
        self.model_dim = (nx, ny, nz)

    def synth_grid(self, kind: int, nx: int, ny: int, nz: int, seed: int = 1729):
        _check(lib().xn_synth_grid_device(self._h, kind, nx, ny, nz, seed))
        self.model_dim = (nx, ny, nz)

    def set_grid_layout(self, mode: int):
        """Residency layout of the grid in HBM (LAYOUT_AUTO / LAYOUT_LINEAR / LAYOUT_BRICKED)."""
        _check(lib().xn_set_grid_layout(self._h, mode))

    def grid_layout(self):
        """(layout resident now, bytes it occupies)."""
        layout, nbytes = C.c_int(), C.c_uint64()
        _check(lib().xn_grid_layout(self._h, C.byref(layout), C.byref(nbytes)))
        return layout.value, nbytes.value

    def download_grid(self) -> Grid:
        nx, ny, nz = self.model_dim
        out = np.empty((nz, ny, nx, 4), dtype=np.uint8)
        _check(lib().xn_download_grid(self._h, out.ctypes.data, out.nbytes))
        return Grid(out)

    def convert_resident_grid(self, chan_diff: int = 0, type: int = TYPE_SPARSE, bind: bool = True,
                              want_nodes: bool = False, std_dev=None):
        """GPU `xenodon convert` of the resident grid -> (Octree | None, stats dict, count, side).
        std_dev selects the --std-dev heuristic (XenodonError with status -5 when the threshold is
        within rounding distance of a cell's deviation: build_octree on the host decides those)."""
        out, count, side, st = C.c_void_p(), C.c_uint64(), C.c_uint64(), BuildStats()
        heur, param = (HEUR_STD_DEV, float(std_dev)) if std_dev is not None else (HEUR_CHAN_DIFF, float(chan_diff))
        _check(lib().xn_convert_resident_grid_ex(self._h, heur, param, type, int(bind),
                                                 C.byref(out) if want_nodes else None, C.byref(count), C.byref(side),
                                                 C.byref(st)))
        tree = None
        if want_nodes:
            try:
                buf = (C.c_char * (count.value * 40)).from_address(out.value)
                tree = Octree(np.frombuffer(buf, dtype=NODE_DTYPE).copy(), side.value)
            finally:
                lib().xn_free(out)
        if bind:
            self.model_dim = (side.value,) * 3
        return tree, {k: getattr(st, k) for k, _ in BuildStats._fields_}, count.value, side.value

    def upload_svo(self, tree: Octree):
        _check(lib().xn_upload_svo(self._h, tree.nodes.ctypes.data, len(tree.nodes), tree.side))
        self.model_dim = (tree.side,) * 3

    def set_target(self, output, display=None):
        display = display or output
        o, d = Rect(*output), Rect(*display)
        _check(lib().xn_set_target(self._h, C.byref(o), C.byref(d)))
        self.output, self.display = tuple(output), tuple(display)

    def set_params(self, voxel_ratio=(1, 1, 1), model_dim=None, emission=1.0):
        md = (C.c_uint32 * 3)(*(model_dim or self.model_dim))
        r = _f3(voxel_ratio)
        _check(lib().xn_set_params(self._h, C.byref(r), C.byref(md), float(emission)))

    def set_precision(self, strict: bool):
        """strict=True: bit-identical to the CPU restatement of the shaders; False (default): fast mode."""
        _check(lib().xn_set_precision(self._h, 1 if strict else 0))

    def set_interleave(self, count: int, index: int):
        _check(lib().xn_set_interleave(self._h, count, index))

    def owned_rays(self) -> int:
        n = C.c_uint64()
        _check(lib().xn_owned_rays(self._h, C.byref(n)))
        return n.value

    def set_target_buffer(self, device_ptr, stride_px: int):
        _check(lib().xn_set_target_buffer(self._h, device_ptr, stride_px))

    def render(self, traversal, camera):
        t = TRAVERSALS[traversal] if isinstance(traversal, str) else traversal
        f, u, p = _f3(camera[0]), _f3(camera[1]), _f3(camera[2])
        _check(lib().xn_render(self._h, t, C.byref(f), C.byref(u), C.byref(p)))

    def render_download_async(self, traversal, camera, host_frame: "PinnedFrame"):
        """Pipelined frame: traversal + asynchronous copy-out into pinned host memory."""
        t = TRAVERSALS[traversal] if isinstance(traversal, str) else traversal
        f, u, p = _f3(camera[0]), _f3(camera[1]), _f3(camera[2])
        _check(lib().xn_render_download_async(self._h, t, C.byref(f), C.byref(u), C.byref(p), host_frame.ptr))

    def render_download_to(self, traversal, camera, host_frame_ptr: int, stride_px: int):
        """Pipelined frame whose owned rows (all, or this context's stripes) land in a shared host frame."""
        t = TRAVERSALS[traversal] if isinstance(traversal, str) else traversal
        f, u, p = _f3(camera[0]), _f3(camera[1]), _f3(camera[2])
        _check(lib().xn_render_download_to(self._h, t, C.byref(f), C.byref(u), C.byref(p), host_frame_ptr, stride_px))

    def signal_after_copy(self, host_flag_ptr: int, value: int):
        _check(lib().xn_signal_after_copy(self._h, host_flag_ptr, value))

    def mark(self, which: int):
        _check(lib().xn_mark(self._h, which))

    def mark_elapsed(self) -> float:
        ms = C.c_double()
        _check(lib().xn_mark_elapsed(self._h, C.byref(ms)))
        return ms.value

    def launch_count(self) -> int:
        n = C.c_uint64()
        _check(lib().xn_launch_count(self._h, C.byref(n)))
        return n.value

    def sync(self) -> float:
        ms = C.c_double()
        _check(lib().xn_sync(self._h, C.byref(ms)))
        return ms.value

    def download(self, out: np.ndarray | None = None) -> np.ndarray:
        w, h = self.output[2], self.output[3]
        if out is None:
            out = np.empty((h, w, 4), dtype=np.uint8)
        _check(lib().xn_download(self._h, out.ctypes.data, 0))
        return out

    def stats_pass(self, traversal, camera, per_ray: bool = True):
        """Instrumented frame -> (steps (h, w) u32 | None, bytes (h, w) u64 | None, (sum_steps, sum_bytes))."""
        t = TRAVERSALS[traversal] if isinstance(traversal, str) else traversal
        f, u, p = _f3(camera[0]), _f3(camera[1]), _f3(camera[2])
        w, h = self.output[2], self.output[3]
        steps = np.empty((h, w), dtype=np.uint32) if per_ray else None
        nbytes = np.empty((h, w), dtype=np.uint64) if per_ray else None
        tot = (C.c_uint64 * 2)()
        _check(lib().xn_render_stats_pass(self._h, t, C.byref(f), C.byref(u), C.byref(p),
                                          steps.ctypes.data if per_ray else None,
                                          nbytes.ctypes.data if per_ray else None, C.byref(tot)))
        return steps, nbytes, (tot[0], tot[1])

    def touch_pass(self, camera, use_skip_table: bool = True):
        """Distinct voxels and distinct 32-byte linear-layout sectors a DDA frame fetches (texture
        residency): the frame's compulsory traffic.  use_skip_table False = every step fetches."""
        f, u, p = _f3(camera[0]), _f3(camera[1]), _f3(camera[2])
        out = (C.c_uint64 * 2)()
        _check(lib().xn_render_touch_pass(self._h, C.byref(f), C.byref(u), C.byref(p), 1 if use_skip_table else 0,
                                          C.byref(out)))
        return int(out[0]), int(out[1])

    # frame buffers shared between processes (one process per GPU)
    def frame_buffer_create(self, w: int, h: int):
        ptr = C.c_void_p()
        handle = (C.c_uint8 * 64)()
        _check(lib().xn_frame_buffer_create(self._h, w, h, C.byref(ptr), handle))
        return ptr.value, bytes(handle)

    def frame_buffer_open(self, handle: bytes) -> int:
        ptr = C.c_void_p()
        buf = (C.c_uint8 * 64)(*handle)
        _check(lib().xn_frame_buffer_open(self._h, buf, C.byref(ptr)))
        return ptr.value

    def frame_buffer_close(self, ptr: int):
        _check(lib().xn_frame_buffer_close(self._h, ptr))

    def frame_buffer_read(self, ptr: int, w: int, h: int) -> np.ndarray:
        out = np.empty((h, w, 4), dtype=np.uint8)
        _check(lib().xn_frame_buffer_read(self._h, ptr, w, h, out.ctypes.data))
        return out


class PinnedFrame:
    """Page-locked host frame (h, w, 4) uint8 for asynchronous copy-out."""

    def __init__(self, w: int, h: int):
        p = C.c_void_p()
        _check(lib().xn_host_alloc(w * h * 4, C.byref(p)))
        self.ptr = p.value
        self.array = np.frombuffer((C.c_uint8 * (w * h * 4)).from_address(self.ptr), dtype=np.uint8).reshape(h, w, 4)

    def free(self):
        if self.ptr:
            self.array = None
            lib().xn_host_free(self.ptr)
            self.ptr = None


def host_register(ptr: int, nbytes: int):
    """Page-lock memory the caller mapped itself (shared memory seen by several processes)."""
    _check(lib().xn_host_register(ptr, nbytes))


def host_unregister(ptr: int):
    _check(lib().xn_host_unregister(ptr))


def _frame_buffer_read_async(self, ptr: int, w: int, h: int, pinned: "PinnedFrame"):
    _check(lib().xn_frame_buffer_read_async(self._h, ptr, w, h, pinned.ptr))


def _copy_sync(self):
    _check(lib().xn_copy_sync(self._h))


Context.frame_buffer_read_async = _frame_buffer_read_async
Context.copy_sync = _copy_sync


@dataclass
class ShaderParameters:
    """reference src/render/RenderContext.h:16-20"""
    voxel_ratio: tuple = (1.0, 1.0, 1.0)
    emission_coeff: float = 1.0


class MultiplexRenderer:
    """Fan-out over the `device {}` entries of a headless configuration
    (reference src/render/MultiplexRenderer.cpp, src/backend/headless/HeadlessDisplay.cpp).

    devices: list of (vkindex, (x, y, w, h)); volume: Grid or Octree (replicated per device).
    """

    def __init__(self, devices, volume, traversal, params: ShaderParameters | None = None):
        params = params or ShaderParameters()
        self.traversal = TRAVERSALS[traversal] if isinstance(traversal, str) else traversal
        is_grid = isinstance(volume, Grid)
        if (self.traversal == DDA) != is_grid:
            name = lib().xn_traversal_name(self.traversal).decode()
            have, need = ("tiff", "svo") if is_grid else ("svo", "tiff")
            raise XenodonError(-1, f"Shader '{name}' is incompatible with model type '{have}' (requires '{need}')")
        self.regions = [tuple(r) for _, r in devices]
        self.display_region = rect_union(self.regions)
        self.contexts = []
        for (index, region) in devices:
            ctx = Context(index)
            if is_grid:
                ctx.upload_grid(volume)
            else:
                ctx.upload_svo(volume)
            ctx.set_target(region, self.display_region)
            ctx.set_params(params.voxel_ratio, None, params.emission_coeff)
            self.contexts.append(ctx)
        self._stats = RenderStats()

    def render(self, camera):
        """One frame on every device; returns after all devices finished (swap_buffers)."""
        for ctx in self.contexts:
            ctx.render(self.traversal, camera)
        st = RenderStats(0, 0, 0.0, 0.0, float("inf"))
        for ctx, region in zip(self.contexts, self.regions):
            ms = ctx.sync()
            st.total_rays += region[2] * region[3]
            st.outputs += 1
            st.total_render_time += ms
            st.max_render_time = max(st.max_render_time, ms)
            st.min_render_time = min(st.min_render_time, ms)
        self._stats = st

    def stats(self) -> RenderStats:
        return self._stats

    def frame(self) -> np.ndarray:
        """Composite of all tiles (HeadlessDisplay::save): (H, W, 4) uint8, background 0xFF000000."""
        _, _, w, h = self.display_region
        out = np.empty((h, w, 4), dtype=np.uint8)
        handles = (C.c_void_p * len(self.contexts))(*[c._h for c in self.contexts])
        enc = Rect()
        _check(lib().xn_frame_gather(handles, len(self.contexts), out.ctypes.data, C.byref(enc)))
        assert (enc.w, enc.h) == (w, h)
        return out

    def close(self):
        for c in self.contexts:
            c.close()
        self.contexts = []
