"""GPU tests of the volume ingest pipeline: xn_upload_grid_tiff reads z slices into
page-locked staging and decodes them on the device.  The resident voxels must equal what the
host reader produces (which is pinned against real libtiff, tests/golden/tiff_golden.npz), for
every sample layout the reader supports, and files it cannot stream (tiles) must take the host
fallback with the same result."""
import os

import numpy as np
import pytest

from util import CAMERAS, write_tiff

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def ingest(xb, path):
    ctx = xb.Context(0)
    try:
        secs = ctx.upload_grid_tiff(path)
        assert secs > 0
        return ctx.model_dim, ctx.download_grid().data
    finally:
        ctx.close()


def test_ingest_matches_libtiff_golden(xb):
    z = np.load(os.path.join(GOLD, "tiff_golden.npz"))
    for name in z.files:
        dims, data = ingest(xb, os.path.join(GOLD, name))
        assert np.array_equal(data, z[name]), name
        assert dims == (data.shape[2], data.shape[1], data.shape[0])


def test_ingest_orientation_tag_matches_libtiff_golden(xb):
    """The device decode applies the Orientation tag as libtiff does (rows / columns reversed)."""
    z = np.load(os.path.join(GOLD, "tiff_orient_golden.npz"))
    for name in z.files:
        assert xb.tiff_stream_info(os.path.join(GOLD, name))["streamable"]
        dims, dev = ingest(xb, os.path.join(GOLD, name))
        assert np.array_equal(dev, z[name]), name


VARIANTS = {
    "rgba_unassociated_strips": dict(spp=4, photometric=2, extra=2, rows_per_strip=3),
    "rgba_associated": dict(spp=4, photometric=2, extra=1),
    "rgba_unspecified_bottom_left": dict(spp=4, photometric=2, orientation=4),
    "rgb_strips": dict(spp=3, photometric=2, rows_per_strip=4),
    "grey_white_is_zero": dict(spp=1, photometric=0),
    "grey_alpha_unassociated": dict(spp=2, photometric=1, extra=2, rows_per_strip=5),
    "rgba_tiled_fallback": dict(spp=4, photometric=2, extra=2, tile=(16, 16)),
}


@pytest.mark.parametrize("name", list(VARIANTS))
def test_ingest_matches_host_reader_on_every_layout(xb, tmp_path, name):
    v = dict(VARIANTS[name])
    spp = v.pop("spp")
    rng = np.random.default_rng(len(name))
    layers = [rng.integers(0, 256, (11, 13, spp), dtype=np.uint8) for _ in range(5)]
    p = tmp_path / (name + ".tif")
    write_tiff(p, layers, **v)
    host = xb.Grid.load_tiff(p).data
    dims, dev = ingest(xb, p)
    assert dims == (13, 11, 5)
    assert np.array_equal(dev, host)
    if name == "rgba_unassociated_strips":  # spot-check the semantics themselves, not only agreement
        a = layers[2][::-1]  # bottom-up rows
        want = (a[..., :3].astype(np.uint32) * a[..., 3:4] + 127) // 255
        assert np.array_equal(dev[2][..., :3], want.astype(np.uint8)) and np.array_equal(dev[2][..., 3], a[..., 3])


def test_ingest_then_render_equals_upload_then_render(xb, tmp_path):
    g = xb.Grid.synthetic(xb.SYNTH_BUNNY, 64, 45, 64).data
    p = tmp_path / "vol.tif"
    xb.Grid(g).save_tiff(p)
    imgs = []
    for how in ("tiff", "host"):
        ctx = xb.Context(0)
        try:
            if how == "tiff":
                ctx.upload_grid_tiff(p)
            else:
                ctx.upload_grid(xb.Grid(g))
            ctx.set_target((0, 0, 160, 90))
            ctx.set_params((1, 1, 1), None, 4.0)
            ctx.render("dda", CAMERAS["orbit"])
            ctx.sync()
            imgs.append(ctx.download())
        finally:
            ctx.close()
    assert np.array_equal(imgs[0], imgs[1]) and imgs[0][..., :3].any()


def test_ingest_errors_are_loud(xb, tmp_path):
    ctx = xb.Context(0)
    try:
        with pytest.raises(xb.XenodonError, match="Failed to open"):
            ctx.upload_grid_tiff(tmp_path / "missing.tif")
        g = np.random.default_rng(0).integers(0, 256, (4, 8, 8, 4), dtype=np.uint8)
        p = tmp_path / "v.tif"
        xb.Grid(g).save_tiff(p, bigtiff=False)
        trunc = tmp_path / "trunc.tif"
        raw = p.read_bytes()
        # keep the directories (written after the pixel data by our writer?) -- cut a middle chunk instead
        trunc.write_bytes(raw[:len(raw) // 2])
        with pytest.raises(xb.XenodonError):
            ctx.upload_grid_tiff(trunc)
        with pytest.raises(xb.XenodonError, match="incompatible|no grid|xn_set_target"):
            ctx.render("dda", CAMERAS["single"])  # a failed ingest leaves no grid behind
    finally:
        ctx.close()
