"""The `xenodon` command line (C++ host over the C ABI): flag set, error texts and `convert`
behave like the reference's (src/main.cpp, src/convert.cpp).  Rendering needs a GPU and is
covered in test_gpu_cli.py; here only what runs on the host."""
import os
import struct
import subprocess

import numpy as np
import pytest

from util import blobby_grid


@pytest.fixture(scope="module")
def cli(xb):
    assert os.path.exists(xb.CLI_PATH)

    def run(*args):
        r = subprocess.run([xb.CLI_PATH, *map(str, args)], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0  # the reference always exits 0 (src/main.cpp:245)
        return r.stdout

    return run


def test_help_and_unknown_subcommand(cli):
    assert "Usage: xenodon <subcommand>" in cli()
    assert "Usage: xenodon <subcommand>" in cli("help")
    assert "--headless <config path>" in cli("help", "render")
    assert "--chan-diff" in cli("help", "convert")
    assert "vkindex" in cli("help", "headless-config")
    assert "Error: No such topic nonsense" in cli("help", "nonsense")
    assert "Error: Invalid subcommand 'frobnicate'" in cli("frobnicate")


def test_render_argument_errors(cli, tmp_path):
    conf = tmp_path / "h.conf"
    conf.write_text("device { vkindex = 0 offset = (0, 0) extent = (64, 36) }\n")
    cases = [
        (["vol.tif"], "Error: Missing required backend --xorg, --headless or --direct"),
        (["--headless", conf, "--xorg", "vol.tif"], "Error: --xorg, --headless and --direct are mutually exclusive"),
        (["--headless", conf], "Error: Missing required positional argument <volume path>"),
        (["--headless", conf, "a.tif", "b.tif"], "Error: Unexpected positional argument 'b.tif'"),
        (["--headless", conf, "--bogus", "a.tif"], "Error: Unrecognized option --bogus"),
        (["--headless", conf, "-e", "1e3", "a.tif"], "Error: Invalid value for <emission coefficient> of parameter -e"),
        (["--headless", conf, "-e", "-1", "a.tif"], "Error: Invalid value for <emission coefficient> of parameter -e"),
        (["--headless", conf, "-r", "1:2", "a.tif"], "Error: Invalid value for <voxel dimension ratio> of parameter -r"),
        (["--headless", conf, "-q", "--quiet", "a.tif"], "Error: Duplicate specification of flag --quiet/-q"),
        (["--headless", conf, "--discard-output", "--output", "x-{}.png", "a.tif"],
         "Error: --dont-save and --output are mutually exclusive"),
        (["--headless"], "Error: Parameter --headless expects argument <config path>"),
        (["--discard-output", "--xorg", "a.tif"], "Error: --dont-save requires --headless"),
    ]
    for args, expected in cases:
        assert expected in cli("render", *args), args


def test_render_backend_errors_are_reported_like_the_reference(cli, tmp_path):
    bad = tmp_path / "bad.conf"
    bad.write_text("device { vkindex = 0 offset = (0, 0) }\n")
    out = cli("render", "--headless", bad, "vol.tif", "--camera", "cam.txt")
    assert "Error: Failed to initialize backend: Configuration error: Missing key 'extent'" in out
    out = cli("render", "--xorg", "vol.tif")
    assert "Error: Failed to initialize backend:" in out


def test_convert_end_to_end_matches_library_and_reference_bytes(cli, xb, tmp_path):
    rng = np.random.default_rng(12)
    g = blobby_grid(rng, 24, 17, 20)
    tif = tmp_path / "v.tif"
    xb.Grid(g).save_tiff(tif)
    for flags, kw in ((["--chan-diff", "0"], dict(chan_diff=0)), (["--rope"], dict(chan_diff=0, type=xb.TYPE_ROPE)),
                      (["--dag", "--chan-diff", "40"], dict(chan_diff=40, type=xb.TYPE_DAG)),
                      (["--std-dev", "10.5"], dict(std_dev=10.5))):
        svo = tmp_path / "o.svo"
        out = cli("convert", *flags, tif, svo)
        assert "Loading source..." in out and "Converting to octree..." in out and "Generated octree:" in out
        assert "24x17x20 = 8160 pixels" in out and " Dimensions: 32x32x32" in out
        tree, st = xb.build_octree(xb.Grid(g), **kw)
        raw = svo.read_bytes()
        assert raw == b"XNDN-SVO" + struct.pack("<QQ", tree.side, len(tree.nodes)) + tree.nodes.tobytes()
        assert f" Total leaves: {st['total_leaves']}" in out and f" Depth: {st['depth']}" in out
        from oracle import xref_model
        if xref_model.available():  # the reference's own convert on the same TIFF gives the same file
            ref = tmp_path / "ref.svo"
            xref_model.convert(tif, ref, **kw)
            assert ref.read_bytes() == raw


def test_convert_argument_errors(cli, tmp_path):
    assert "Error: --dag and --rope are mutually exclusive" in cli("convert", "--dag", "--rope", "a.tif", "b.svo")
    assert "Error: --std-dev and --chan-diff are mutually exclusive" in cli(
        "convert", "--std-dev", "1", "--chan-diff", "2", "a.tif", "b.svo")
    assert "Error: Invalid value for <channel difference> of parameter --chan-diff" in cli(
        "convert", "--chan-diff", "256", "a.tif", "b.svo")
    assert "Error: Missing required positional argument <destination svo path>" in cli("convert", "a.tif")
    assert "Error reading '" in cli("convert", tmp_path / "missing.tif", tmp_path / "o.svo")
