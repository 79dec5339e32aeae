"""Host-side logic of the multi-GPU path: how one frame is partitioned over ranks and how the
tiles come back to rank 0.  Pure Python + torch.distributed (any backend), so it is testable
with gloo on CPU; the GPU work itself is in the C ABI (xn_set_interleave,
xn_set_target_buffer, xn_frame_buffer_*).

Partition = the reference's sort-first screen split (one `device {}` rectangle per GPU, uv from
the global display rectangle so tiles are seamless; reference src/render/Renderer.cpp:236-268),
in two flavours:
  * stripes: 16-row stripes dealt round-robin (balanced; what xn_set_interleave renders);
  * bands  : one contiguous horizontal band per rank (what a 1-block-per-GPU headless.conf says).
"""
from __future__ import annotations

import math
import os

import numpy as np

STRIPE_ROWS = 16  # = BLOCK_H of the traversal kernels


def frame_for(base, n_gpus: int, weak: bool = True):
    """Frame size with ~n_gpus times the rays of `base` (w, h) at the same aspect ratio."""
    w, h = base
    if not weak or n_gpus == 1:
        return w, h
    hh = int(round(h * math.sqrt(n_gpus) / STRIPE_ROWS)) * STRIPE_ROWS
    ww = int(round(hh * w / h))
    return ww, hh


def stripe_rows(height: int, count: int, index: int):
    """[(y0, rows)] of the 16-row stripes rank `index` of `count` owns (xn_set_interleave)."""
    if count <= 0 or not 0 <= index < count:
        raise ValueError("interleave index must be < count")
    out = []
    s = index
    while s * STRIPE_ROWS < height:
        y0 = s * STRIPE_ROWS
        out.append((y0, min(STRIPE_ROWS, height - y0)))
        s += count
    return out


def owned_rays(width: int, height: int, count: int, index: int) -> int:
    return width * sum(r for _, r in stripe_rows(height, count, index))


def band_rows(height: int, n: int):
    """Row boundaries [y_0 .. y_n] of n contiguous bands, aligned to the 16-row stripe grid."""
    edges = [((height * r) // n // STRIPE_ROWS) * STRIPE_ROWS for r in range(n)] + [height]
    for a, b in zip(edges, edges[1:]):
        if b <= a:
            raise ValueError("frame too small for this many bands")
    return edges


def combine_stats(per_rank):
    """RenderStats::combine over ranks (reference src/render/RenderStats.cpp:13-20):
    per_rank = [(rays, ms)] -> dict with the reference's summed-time rate and the aggregate rate."""
    rays = sum(r for r, _ in per_rank)
    total_ms = sum(ms for _, ms in per_rank)
    max_ms = max(ms for _, ms in per_rank)
    return {
        "total_rays": rays, "outputs": len(per_rank), "total_render_time": total_ms,
        "max_render_time": max_ms, "min_render_time": min(ms for _, ms in per_rank),
        "mrays_per_s_reference": rays / (total_ms * 1000.0),  # rays / summed device time
        "mrays_per_s_aggregate": rays / (max_ms * 1000.0),    # rays / slowest device
    }


def gather_stripes(dist, local_rows, width: int, height: int, rank: int, world: int):
    """Collects row-interleaved tiles on rank 0 with any torch.distributed backend.

    local_rows: int32 tensor (owned_row_count, width) holding this rank's stripes top to bottom.
    Returns the (height, width) frame on rank 0, None elsewhere.
    """
    import torch

    counts = [sum(r for _, r in stripe_rows(height, world, r_)) for r_ in range(world)]
    if rank == 0:
        frame = torch.empty((height, width), dtype=local_rows.dtype)
        bufs = [torch.empty((c, width), dtype=local_rows.dtype) for c in counts]
        bufs[0].copy_(local_rows)
        reqs = [dist.irecv(bufs[r_], src=r_) for r_ in range(1, world)]
        for q in reqs:
            q.wait()
        for r_ in range(world):
            off = 0
            for (y0, rows) in stripe_rows(height, world, r_):
                frame[y0:y0 + rows] = bufs[r_][off:off + rows]
                off += rows
        return frame
    dist.send(local_rows.contiguous(), dst=0)
    return None


# --------------------------------------------------------------------------------------------
# One host frame shared by N processes (one per GPU): every rank delivers its own stripes over its
# own PCIe link (xn_render_download_to) and flags completion in stream order
# (xn_signal_after_copy); the consumer polls the flags.  Replaces the host composite of
# HeadlessDisplay::save (reference src/backend/headless/HeadlessDisplay.cpp:59-76) without a
# gathering device.  Nothing here touches CUDA: page-locking is an injected callable, so the
# protocol runs under gloo on CPU (tests/test_multi_gloo.py).
# --------------------------------------------------------------------------------------------
def interleave_pages(addr: int, size: int) -> str:
    """mbind(MPOL_INTERLEAVE) over the online NUMA nodes for a not-yet-touched mapping, so that N
    GPUs writing one segment do not all land on the creating rank's memory controllers; returns
    what happened as text.  XN_SHM_INTERLEAVE=0 leaves the default policy."""
    if os.environ.get("XN_SHM_INTERLEAVE", "1") == "0":
        return "default policy (XN_SHM_INTERLEAVE=0)"
    try:
        import ctypes
        ids = []
        for part in open("/sys/devices/system/node/online").read().strip().split(","):
            a, _, b = part.partition("-")
            ids += list(range(int(a), int(b or a) + 1))
        if len(ids) < 2:
            return "one NUMA node"
        mask = 0
        for i in ids:
            mask |= 1 << i
        libc = ctypes.CDLL(None, use_errno=True)
        words = (max(ids) + 64) // 64
        nodemask = (ctypes.c_ulong * words)(*[(mask >> (64 * w)) & (2**64 - 1) for w in range(words)])
        a0 = addr & ~4095
        r = libc.syscall(237, ctypes.c_void_p(a0), ctypes.c_ulong(size + addr - a0), ctypes.c_int(3), nodemask,
                         ctypes.c_ulong(64 * words + 1), ctypes.c_uint(0))  # SYS_mbind, MPOL_INTERLEAVE
        return f"pages interleaved over {len(ids)} NUMA nodes" if r == 0 else f"default policy (mbind errno {ctypes.get_errno()})"
    except Exception as e:  # not Linux / no sysfs / no syscall: the default policy stays
        return f"default policy ({type(e).__name__})"


class SharedHostFrames:
    """`count` frames of w*h RGBA8 behind one flag page, in ONE POSIX shared-memory segment mapped by
    every rank.  `register(addr, bytes)` page-locks the mapping for CUDA (xn_host_register) so that
    device-to-host copies into it are asynchronous; None on a CPU-only run.
    flags[rank, slot] = sequence number of the last frame `rank` delivered into `slot`;
    ack[0] = last frame the consumer has released."""

    FLAG_BYTES = 4096
    ACK_OFFSET = 2048  # ranks * count * 4 bytes of flags must stay below this

    def __init__(self, name, w, h, count, n_ranks, create, register=None, unregister=None):
        from multiprocessing import shared_memory
        if n_ranks * count * 4 > self.ACK_OFFSET:
            raise ValueError("too many ranks x slots for the flag page")
        self.frame_bytes = w * h * 4
        self.flag_bytes = self.FLAG_BYTES
        self.count = count
        size = self.flag_bytes + count * self.frame_bytes
        self.shm = shared_memory.SharedMemory(name=name, create=create, size=size)
        if not create:
            # only the creating rank owns the segment: keep this process's resource tracker from
            # unlinking (and warning about) a segment it merely attached to
            try:
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        self.buf = np.frombuffer(self.shm.buf, dtype=np.uint8)
        self.base = self.buf.ctypes.data
        # before anything touches the pages (pinning allocates them)
        self.numa = interleave_pages(self.base, size) if create else None
        self._unregister = unregister
        if register is not None:
            register(self.base, size)
        self.flags = self.buf[:n_ranks * count * 4].view(np.uint32).reshape(n_ranks, count)
        self.ack = self.buf[self.ACK_OFFSET:self.ACK_OFFSET + 4].view(np.uint32)
        self.frames = [self.buf[self.flag_bytes + i * self.frame_bytes:self.flag_bytes + (i + 1) * self.frame_bytes]
                       .reshape(h, w, 4) for i in range(count)]
        self.create = create
        if create:
            self.flags[...] = 0
            self.ack[0] = 0

    def frame_ptr(self, i):
        return self.base + self.flag_bytes + i * self.frame_bytes

    def flag_ptr(self, rank, slot):
        return self.base + (rank * self.flags.shape[1] + slot) * 4

    def close(self):
        try:
            if self._unregister is not None:
                self._unregister(self.base)
        except Exception:
            pass
        self.flags = self.ack = self.frames = self.buf = None
        try:
            self.shm.close()
            if self.create:
                self.shm.unlink()
        except Exception:
            pass


class HostFrameRing:
    """Producer / consumer order over SharedHostFrames.  Frames are numbered 1, 2, ...; frame seq
    lives in slot seq % count.  A producer may fill a slot only after the consumer has released the
    frame that slot held before (seq - count); the consumer sees frame seq complete when every
    rank's flag for its slot has reached seq.  Busy-waiting on purpose: the waits are a fraction of a
    frame time and the flags are written by CUDA host callbacks, not by Python."""

    def __init__(self, shared: SharedHostFrames, rank: int):
        self.s, self.rank = shared, rank

    def slot(self, seq: int) -> int:
        return seq % self.s.count

    def acquire(self, seq: int) -> int:
        """Producer: wait until slot(seq) may be overwritten; returns the slot."""
        need = seq - self.s.count
        while need > 0 and int(self.s.ack[0]) < need:
            pass
        return self.slot(seq)

    def wait_complete(self, seq: int) -> int:
        """Consumer: wait until every rank has delivered frame seq; returns the slot."""
        fl = self.s.flags[:, self.slot(seq)]
        while int(fl.min()) < seq:
            pass
        return self.slot(seq)

    def release(self, seq: int):
        """Consumer: frame seq (and every earlier one) is no longer needed."""
        self.s.ack[0] = seq
