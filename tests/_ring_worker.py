"""Worker of tests/test_multi_gloo.py::test_shared_host_frame_ring (torch.distributed.run, gloo, CPU).
Every rank fills its own 16-row stripes of a frame that lives in POSIX shared memory -- numpy stores
stand in for the device-to-host copies -- and flags completion; rank 0 consumes two frames behind and
checks that each frame is complete and not yet overwritten when it is released (the protocol of
bench.py's multi-GPU e2e region, xenodon_b200.distributed.HostFrameRing)."""
import json
import os
import random
import sys
import time

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from xenodon_b200 import distributed as xd  # noqa: E402


def value(seq, y):
    return (seq * 2654435761 + y * 40503) & 0xFFFFFFFF


def main():
    out_path = sys.argv[1]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    W, H, SLOTS, FRAMES = 64, 100, 3, 60  # the last stripe is short
    name = [f"xn_ring_test_{os.getpid()}" if rank == 0 else None]
    dist.broadcast_object_list(name, src=0)
    shared = xd.SharedHostFrames(name[0], W, H, SLOTS, world, create=True) if rank == 0 else None
    dist.barrier()
    if rank != 0:
        shared = xd.SharedHostFrames(name[0], W, H, SLOTS, world, create=False)
    dist.barrier()
    ring = xd.HostFrameRing(shared, rank)
    rng = random.Random(rank)
    mine = xd.stripe_rows(H, world, rank)
    bad = []

    def produce(seq):
        slot = ring.acquire(seq)
        px = shared.frames[slot].view(np.uint32).reshape(H, W)
        for (y0, n) in mine:
            for y in range(y0, y0 + n):
                px[y, :] = value(seq, y)
        time.sleep(rng.random() * 0.002)  # jitter: the ranks finish in varying order
        shared.flags[rank, slot] = seq

    def consume(seq):
        slot = ring.wait_complete(seq)
        px = shared.frames[slot].view(np.uint32).reshape(H, W)
        want = np.array([value(seq, y) for y in range(H)], dtype=np.uint32)[:, None]
        if not np.array_equal(px, np.broadcast_to(want, (H, W))):
            bad.append(seq)
        ring.release(seq)

    lag = SLOTS - 1
    for i in range(FRAMES):
        produce(1 + i)
        if rank == 0 and i >= lag:
            consume(1 + i - lag)
    if rank == 0:
        for seq in range(FRAMES + 1 - lag, FRAMES + 1):
            consume(seq)
        with open(out_path, "w") as f:
            json.dump({"world": world, "frames": FRAMES, "bad": bad, "released": int(shared.ack[0]),
                       "numa": shared.numa}, f)
    dist.barrier()
    shared.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
