"""N > 1 host logic on CPU: partition arithmetic, and a world_size-2 gloo run in which the
stripes shaded by two ranks are gathered on rank 0 and must equal the single-rank frame."""
import json
import os
import subprocess
import sys

import pytest

from xenodon_b200 import distributed as xd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_stripe_partition_covers_the_frame_exactly_once():
    for h in (1, 15, 16, 17, 70, 1080, 1520, 3056):
        for n in (1, 2, 3, 4, 8):
            seen = [0] * h
            for r in range(n):
                for y0, rows in xd.stripe_rows(h, n, r):
                    assert y0 % 16 == 0 and 0 < rows <= 16
                    for y in range(y0, y0 + rows):
                        seen[y] += 1
            assert seen == [1] * h
            assert sum(xd.owned_rays(100, h, n, r) for r in range(n)) == 100 * h
    with pytest.raises(ValueError):
        xd.stripe_rows(100, 2, 2)


def test_band_partition_and_weak_frames():
    assert xd.band_rows(1080, 1) == [0, 1080]
    assert xd.band_rows(1520, 2) == [0, 752, 1520]
    e = xd.band_rows(3056, 8)
    assert e[0] == 0 and e[-1] == 3056 and all(b > a and a % 16 == 0 for a, b in zip(e, e[1:]))
    assert xd.frame_for((1920, 1080), 1) == (1920, 1080)
    assert xd.frame_for((1920, 1080), 4) == (3840, 2160)
    for n in (2, 8):
        w, h = xd.frame_for((1920, 1080), n)
        assert h % 16 == 0 and abs(w * h / (1920 * 1080) - n) < 0.03 * n
    assert xd.frame_for((1920, 1080), 8, weak=False) == (1920, 1080)


def test_stats_combine_like_the_reference():
    st = xd.combine_stats([(1000, 2.0), (3000, 4.0)])
    assert st["total_rays"] == 4000 and st["outputs"] == 2 and st["total_render_time"] == 6.0
    assert st["max_render_time"] == 4.0 and st["min_render_time"] == 2.0
    assert abs(st["mrays_per_s_reference"] - 4000 / 6000.0) < 1e-12  # rays / SUMMED device time
    assert abs(st["mrays_per_s_aggregate"] - 1.0) < 1e-12


def test_two_rank_gloo_gather_equals_single_rank_frame(tmp_path, xo):
    out = tmp_path / "result.json"
    env = dict(os.environ, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29631", os.path.join(ROOT, "tests", "_gloo_worker.py"), str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    res = json.loads(out.read_text())
    assert res["world"] == 2 and res["equal"] is True
    assert res["stats"]["total_rays"] == 100 * 70 and res["stats"]["max_render_time"] == 2.0


def test_shared_host_frame_ring(tmp_path):
    """The multi-GPU e2e protocol on CPU: two gloo ranks deliver their stripes into one shared host
    frame ring (three slots), rank 0 consumes two frames behind; every frame must be complete when
    it is seen and intact until it is released."""
    out = tmp_path / "ring.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29633", os.path.join(ROOT, "tests", "_ring_worker.py"), str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, OMP_NUM_THREADS="1"), cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    res = json.loads(out.read_text())
    assert res["world"] == 2 and res["bad"] == [] and res["released"] == res["frames"]


def test_host_frame_ring_rejects_too_many_flags():
    from xenodon_b200 import distributed as xd
    with pytest.raises(ValueError):
        xd.SharedHostFrames("xn_never_created", 8, 8, 64, 16, create=True)
