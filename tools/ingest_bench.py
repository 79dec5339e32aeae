#!/usr/bin/env python3
"""Volume ingest: wall time of TIFF -> resident grid, old path (host decode of the
whole stack + one bulk upload, what the reference's Grid::load_tiff + DdaRaytraceResources do) vs
the pipelined xn_upload_grid_tiff.  usage: python tools/ingest_bench.py [nx ny nz]"""
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import xenodon_b200 as xb  # noqa: E402

nx, ny, nz = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (1024, 1024, 512)
path = os.path.join(tempfile.gettempdir(), f"xn_ingest_{nx}x{ny}x{nz}.tif")
ctx = xb.Context(0)
ctx.set_grid_layout(xb.LAYOUT_LINEAR)
ctx.synth_grid(xb.SYNTH_TNG, nx, ny, nz)
vol = ctx.download_grid()
t0 = time.perf_counter()
vol.save_tiff(path)
t_write = time.perf_counter() - t0
gib = vol.data.nbytes / 2**30
out = {"volume": f"{nx}x{ny}x{nz} RGBA8 ({gib:.2f} GiB), BigTIFF, one strip per slice, page cache warm",
       "threads": os.cpu_count()}
for rep in range(2):
    t0 = time.perf_counter()
    host = xb.Grid.load_tiff(path)
    t_read = time.perf_counter() - t0
    t0 = time.perf_counter()
    ctx.upload_grid(host)
    t_up = time.perf_counter() - t0
    del host
    t_pipe = ctx.upload_grid_tiff(path)
    same = np.array_equal(ctx.download_grid().data, vol.data)
    out[f"run{rep}"] = {"host_decode_s": round(t_read, 3), "bulk_upload_s": round(t_up, 3),
                        "old_path_GiB_s": round(gib / (t_read + t_up), 3), "pipelined_s": round(t_pipe, 3),
                        "pipelined_GiB_s": round(gib / t_pipe, 3), "identical": bool(same)}
ctx.close()
os.unlink(path)
print(json.dumps(out))
