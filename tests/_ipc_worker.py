"""Worker of tests/test_gpu_multi.py: one process per rank (gloo rendezvous), every rank shades
its 16-row stripes and stores them straight into rank 0's frame buffer through a CUDA IPC
mapping (the multi-GPU data path of bench.py; ranks may share one physical GPU)."""
import json
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import xenodon_b200 as xb  # noqa: E402
from oracle import xo  # noqa: E402
from util import CAMERAS, blobby_grid  # noqa: E402
from xenodon_b200 import distributed as xd  # noqa: E402


def main():
    out_path, traversal = sys.argv[1], sys.argv[2]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    device = rank % xb.device_count()
    W, H = 200, 120
    grid = blobby_grid(np.random.default_rng(77), 40, 40, 40)
    ctx = xb.Context(device)
    ctx.set_precision(True)
    if traversal == "dda":
        ctx.upload_grid(xb.Grid(grid))
        tree = None
    else:
        tree, _ = xb.build_octree(xb.Grid(grid), chan_diff=0, type=xb.TYPE_ROPE)
        ctx.upload_svo(tree)
    ctx.set_target((0, 0, W, H))
    ctx.set_params((1, 1, 1), None, 3.0)
    ctx.set_interleave(world, rank)
    assert ctx.owned_rays() == xd.owned_rays(W, H, world, rank)
    if rank == 0:
        ptr, handle = ctx.frame_buffer_create(W, H)
        obj = [handle]
    else:
        obj = [None]
    dist.broadcast_object_list(obj, src=0)
    if rank != 0:
        ptr = ctx.frame_buffer_open(obj[0])
    ctx.set_target_buffer(ptr, W)
    cam = CAMERAS["orbit"]
    ctx.render(traversal, cam)
    ms = ctx.sync()
    dist.barrier()  # every rank's stores have landed
    if rank == 0:
        frame = ctx.frame_buffer_read(ptr, W, H)
        kw = dict(camera=cam, output=(0, 0, W, H), emission=3.0, want_stats=False)
        ref = (xo.render("dda", grid=grid, **kw) if traversal == "dda"
               else xo.render(traversal, nodes=tree.nodes, side=tree.side, **kw))[0]
        with open(out_path, "w") as f:
            json.dump({"equal": bool(np.array_equal(frame, ref)), "world": world, "ms": ms,
                       "nonzero": int(frame[..., :3].any(axis=-1).sum())}, f)
    dist.barrier()
    if rank != 0:
        ctx.frame_buffer_close(ptr)
    dist.barrier()
    if rank == 0:
        ctx.frame_buffer_close(ptr)
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
