"""The multi-GPU data path on whatever GPUs are present (ranks share GPU 0 when there is only
one): row-interleaved partition + kernel stores into rank 0's frame through CUDA IPC."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from util import CAMERAS, blobby_grid

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("traversal,world", [("dda", 2), ("esvo", 3)])
def test_ranks_store_their_stripes_into_rank0_frame(tmp_path, traversal, world):
    out = tmp_path / "r.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29641", os.path.join(ROOT, "tests", "_ipc_worker.py"),
           str(out), traversal]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT,
                       env=dict(os.environ, OMP_NUM_THREADS="2"))
    assert r.returncode == 0, r.stderr[-3000:]
    res = json.loads(out.read_text())
    assert res["world"] == world and res["equal"] is True and res["nonzero"] > 1000


def test_interleaved_contexts_compose_the_full_frame(xb, xo):
    """In-process: three contexts with interleave (3, i) writing into ONE shared target buffer."""
    g = blobby_grid(np.random.default_rng(78), 32, 32, 32)
    W, H = 150, 100
    ctxs = [xb.Context(0) for _ in range(3)]
    ptr, _ = ctxs[0].frame_buffer_create(W, H)
    total = 0
    for i, c in enumerate(ctxs):
        c.set_precision(True)
        c.upload_grid(xb.Grid(g))
        c.set_target((0, 0, W, H))
        c.set_params((1, 1, 1), None, 2.0)
        c.set_interleave(3, i)
        c.set_target_buffer(ptr, W)
        total += c.owned_rays()
        c.render("dda", CAMERAS["orbit"])
    for c in ctxs:
        c.sync()
    assert total == W * H
    frame = ctxs[0].frame_buffer_read(ptr, W, H)
    ref = xo.render("dda", grid=g, camera=CAMERAS["orbit"], output=(0, 0, W, H), emission=2.0, want_stats=False)[0]
    assert np.array_equal(frame, ref)
    steps, _, (tot_steps, _) = ctxs[1].stats_pass("dda", CAMERAS["orbit"])
    rsteps = xo.render("dda", grid=g, camera=CAMERAS["orbit"], output=(0, 0, W, H), emission=2.0)[1]
    own = np.zeros(H, bool)
    for s in range(1, (H + 15) // 16, 3):
        own[s * 16:(s + 1) * 16] = True
    assert np.array_equal(steps[own], rsteps[own]) and not steps[~own].any()
    assert tot_steps == int(rsteps[own].sum())
    ctxs[0].frame_buffer_close(ptr)
    for c in ctxs:
        c.close()


def test_pipelined_frame_output_and_async_frame_read(xb, xo):
    """xn_render_download_async (three device targets in rotation + copy stream) and
    xn_frame_buffer_read_async deliver the same pixels as the blocking calls."""
    g = blobby_grid(np.random.default_rng(79), 32, 32, 32)
    W, H = 160, 96
    ctx = xb.Context(0)
    ctx.set_precision(True)
    ctx.upload_grid(xb.Grid(g))
    ctx.set_target((0, 0, W, H))
    ctx.set_params((1, 1, 1), None, 2.0)
    cams = [CAMERAS["orbit"], CAMERAS["single"], CAMERAS["inside"], CAMERAS["axis_neg"], CAMERAS["orbit"]]
    frames = [xb.PinnedFrame(W, H) for _ in cams]
    for cam, fr in zip(cams, frames):  # five frames in flight through three targets: the rotation wraps
        ctx.render_download_async("dda", cam, fr)
    ms = ctx.sync()
    assert ms > 0 and ctx.launch_count() == len(cams)
    for cam, fr in zip(cams, frames):
        ref = xo.render("dda", grid=g, camera=cam, output=(0, 0, W, H), emission=2.0, want_stats=False)[0]
        assert np.array_equal(fr.array, ref)
    # async read of a shared frame buffer
    ptr, _ = ctx.frame_buffer_create(W, H)
    ctx.set_target_buffer(ptr, W)
    ctx.mark(0)
    ctx.render("dda", cams[0])
    ctx.mark(1)
    ctx.frame_buffer_read_async(ptr, W, H, frames[1])
    ctx.copy_sync()
    assert ctx.mark_elapsed() > 0
    assert np.array_equal(frames[1].array, frames[0].array)
    with pytest.raises(xb.XenodonError, match="external buffer"):
        ctx.render_download_async("dda", cams[0], frames[2])
    ctx.set_target_buffer(None, 0)
    ctx.frame_buffer_close(ptr)
    for fr in frames:
        fr.free()
    ctx.close()
