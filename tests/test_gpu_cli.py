"""End to end through the drop-in CLI on a GPU: `xenodon convert` + `xenodon render --headless`
with the reference's command lines; saved PNGs, stats file and log lines are checked against
the oracle / the reference's formats."""
import os
import subprocess

import numpy as np
import pytest

from util import blobby_grid

pytestmark = pytest.mark.gpu


def run(xb, *args):
    r = subprocess.run([xb.CLI_PATH, *map(str, args)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0
    return r.stdout + r.stderr


def test_render_headless_tiff_and_svo(xb, xo, tmp_path):
    from PIL import Image
    from xenodon_b200 import cameras
    rng = np.random.default_rng(31)
    g = blobby_grid(rng, 32, 23, 32)
    tif = tmp_path / "vol.tif"
    xb.Grid(g).save_tiff(tif)
    conf = tmp_path / "headless.conf"
    # two `device {}` blocks on the same GPU: exercises the multi-device path on one device
    conf.write_text("device {\n    vkindex = 0\n    offset = (0, 0)\n    extent = (96, 108)\n}\n"
                    "device {\n    vkindex = 0\n    offset = (96, 0)\n    extent = (96, 108)\n}\n")
    cam = tmp_path / "cam.txt"
    frames = cameras.camera_rotate(5)
    cam.write_text(cameras.to_text(frames))
    stats = tmp_path / "stats.txt"
    out = run(xb, "render", "--headless", conf, tif, "--camera", cam, "-e", "4", "--output",
              tmp_path / "out-{:0>3}.png", "--stats-output", stats)
    for line in ("Setup: 2 devices, with 1, 1 outputs", "Model file type: 'tiff'", "Using shader 'dda'",
                 "32x23x32 = 23552 pixels", "Model dimensions: 32x23x32", "Total resolution: 192x108 pixels",
                 "Starting render loop...", "Saving frame 4...", "Saved stats to"):
        assert line in out, line
    for i in range(5):
        img = np.asarray(Image.open(tmp_path / f"out-{i:03d}.png"))
        c = frames[i]
        ref = xo.render("dda", grid=g, camera=(tuple(c[0]), tuple(c[1]), tuple(c[2])), output=(0, 0, 192, 108),
                        emission=4.0, want_stats=False)[0]
        assert np.abs(img.astype(int) - ref.astype(int)).max() <= 1  # CLI runs the default (fast) mode
    lines = stats.read_text().splitlines()
    assert lines[0] == f"total rays: {5 * 192 * 108}" and lines[4] == "frames: 5"
    assert lines[6].startswith("frame 0: 20736 rays, 2, ") and lines[6].endswith(" mray/s")

    # convert + every SVO traversal, --repeat, --discard-output
    svo, rope = tmp_path / "vol.svo", tmp_path / "vol-rope.svo"
    assert "Built on the GPU" in run(xb, "convert", tif, svo)
    assert "Built on the GPU" in run(xb, "convert", "--rope", tif, rope)
    # the host builder writes the same bytes
    for flags, path in (([], svo), (["--rope"], rope)):
        host = tmp_path / "host.svo"
        assert "Built on the host" in run(xb, "convert", "--host", *flags, tif, host)
        assert host.read_bytes() == path.read_bytes()
    assert "Built on the GPU" in run(xb, "convert", "--dag", tif, tmp_path / "dag.svo")
    assert "Built on the host" in run(xb, "convert", "--host", "--dag", tif, tmp_path / "dag_host.svo")
    assert (tmp_path / "dag.svo").read_bytes() == (tmp_path / "dag_host.svo").read_bytes()
    for shader, path in (("svo-naive", svo), ("svo-df", svo), ("esvo", svo), ("svo-rope", rope)):
        out = run(xb, "render", "--headless", conf, path, "-s", shader, "--camera", cam, "--repeat", "2",
                  "--discard-output", "--stats-output", stats, "-e", "4")
        assert f"Using shader '{shader}'" in out and "Model dimensions: 32x32x32" in out
        assert stats.read_text().splitlines()[4] == "frames: 10"  # 5 camera lines x repeat 2
    out = run(xb, "render", "--headless", conf, svo, "-s", "dda", "--camera", cam, "--discard-output")
    assert "Error: Shader 'dda' is incompatible with model type 'svo' (requires 'tiff')" in out
    out = run(xb, "render", "--headless", conf, svo, "--volume-type", "tiff", "--camera", cam, "--discard-output")
    assert "Error: Failed to open" in out
    assert "GPU 0:" in run(xb, "sysinfo")
