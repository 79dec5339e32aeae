#!/usr/bin/env python3
"""the same volume as a sparse tree and as a DAG (`xenodon convert --dag`, shared
subtrees), traversed by the kernels that work on DAGs (all but svo_rope).  Reports node counts,
resident bytes of the compact residency, and Mrays/s over camera.txt frames."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import xenodon_b200 as xb  # noqa: E402
from xenodon_b200 import cameras  # noqa: E402

W, H = 1920, 1080
cams = cameras.camera_benchmark()
frames = list(range(5, 150, 5))
out = {"volume": "bunny 512x361x512 seed 1729", "frame": f"{W}x{H}", "frames": len(frames)}
grid = xb.Grid.synthetic(xb.SYNTH_BUNNY, 512, 361, 512)
for chan_diff in (0, 8):
    for name, typ in (("sparse", xb.TYPE_SPARSE), ("dag", xb.TYPE_DAG)):
        t0 = time.perf_counter()
        tree, st = xb.build_octree(grid, chan_diff=chan_diff, type=typ)
        build_s = time.perf_counter() - t0
        internal = int((tree.nodes["is_leaf_depth"] >> 31 == 0).sum())
        ctx = xb.Context(0)
        ctx.upload_svo(tree)
        ctx.set_target((0, 0, W, H))
        ctx.set_params((1, 1, 1), (tree.side,) * 3, 10.0)
        row = {"nodes": len(tree.nodes), "internal_nodes": internal, "compact_MiB": round(internal * 32 / 2**20, 1),
               "host_build_s": round(build_s, 1)}
        for trav in ("esvo", "svo-naive", "svo-df"):
            for i in frames[:3]:
                ctx.render(trav, tuple(map(tuple, cams[i])))
                ctx.sync()
            ms = 0.0
            for i in frames:
                ctx.render(trav, tuple(map(tuple, cams[i])))
                ms += ctx.sync()
            row[trav] = round(W * H * len(frames) / (ms / 1e3) / 1e6, 1)
        ctx.close()
        out[f"chan_diff_{chan_diff}_{name}"] = row
print(json.dumps(out))
