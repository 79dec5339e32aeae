"""Worker of tests/test_multi_gloo.py (launched by torch.distributed.run, gloo backend, CPU).
Each rank shades the 16-row stripes it owns -- with the CPU oracle standing in for the GPU
kernel -- and the tiles are gathered on rank 0 exactly as the multi-GPU path does."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import xo  # noqa: E402
from util import CAMERAS, blobby_grid  # noqa: E402
from xenodon_b200 import distributed as xd  # noqa: E402


def main():
    out_path = sys.argv[1]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    W, H = 100, 70  # not a multiple of 16: the last stripe is short
    grid = blobby_grid(np.random.default_rng(21), 24, 24, 24)
    cam = CAMERAS["orbit"]
    rows = []
    for (y0, n) in xd.stripe_rows(H, world, rank):
        img = xo.render("dda", grid=grid, camera=cam, output=(0, y0, W, n), display=(0, 0, W, H), emission=2.0,
                        threads=1, want_stats=False)[0]
        rows.append(img.view(np.uint32).reshape(n, W).astype(np.int32))
    local = torch.from_numpy(np.concatenate(rows, axis=0))
    assert local.shape[0] * W == xd.owned_rays(W, H, world, rank)
    frame = xd.gather_stripes(dist, local, W, H, rank, world)
    # stats combine: every rank contributes (rays, ms)
    mine = torch.tensor([float(local.shape[0] * W), 1.0 + rank], dtype=torch.float64)
    allst = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(allst, mine)
    if rank == 0:
        full = xo.render("dda", grid=grid, camera=cam, output=(0, 0, W, H), emission=2.0, threads=1,
                         want_stats=False)[0].view(np.uint32).reshape(H, W).astype(np.int32)
        st = xd.combine_stats([(int(a[0]), float(a[1])) for a in allst])
        with open(out_path, "w") as f:
            json.dump({"equal": bool(np.array_equal(frame.numpy(), full)), "world": world, "stats": st}, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
