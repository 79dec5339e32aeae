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
