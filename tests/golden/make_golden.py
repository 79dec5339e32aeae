#!/usr/bin/env python3
"""Regenerates tests/golden/*.npz / *.tif from THE REFERENCE ITSELF, in the build container.

Needs /root/reference and `make -C oracle ref` (oracle/_ref/libxnref_glsl.so = the reference's
own shader text compiled as C++; oracle/_ref/libxnref_model.so = its own src/model code with
the real libtiff).  The outputs are small and committed, so the GPU box -- which has neither
the reference nor oracle/_ref built from it -- can still check against reference outputs.

    python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import xref, xref_model  # noqa: E402
from util import CAMERAS, blobby_grid, random_grid  # noqa: E402

TRAVERSALS = ["dda", "svo-naive", "svo-df", "esvo", "svo-rope"]


def render_cases():
    """(name, grid, camera, output, display, ratio, emission)"""
    rng = np.random.default_rng(20191)
    g1 = random_grid(rng, 16, 16, 16)
    g2 = blobby_grid(rng, 40, 29, 33)
    return [
        ("rand16_single", g1, CAMERAS["single"], (0, 0, 96, 54), (0, 0, 96, 54), (1, 1, 1), 2.0),
        ("rand16_inside", g1, CAMERAS["inside"], (0, 0, 96, 54), (0, 0, 96, 54), (1, 1, 1), 1.0),
        ("blob_orbit", g2, CAMERAS["orbit"], (0, 0, 96, 54), (0, 0, 96, 54), (1, 1, 1), 3.0),
        ("blob_oblique_aniso_tile", g2, CAMERAS["aniso"], (13, 7, 70, 40), (0, 0, 120, 64), (1.0, 2.0, 0.5), 4.0),
        ("blob_axis_neg", g2, CAMERAS["axis_neg"], (0, 0, 96, 54), (0, 0, 96, 54), (1, 1, 1), 2.5),
    ]


def orientation_fixtures():
    """Every value of the Orientation tag (274) through the real libtiff (Grid::load_tiff ->
    TIFFReadRGBAImage): which of them reverse rows, columns, both or neither."""
    from util import write_tiff
    rng = np.random.default_rng(274)
    layers = [rng.integers(0, 256, (4, 5, 3), dtype=np.uint8) for _ in range(2)]
    out = {}
    for o in range(1, 9):
        name = f"orient_{o}.tif"
        write_tiff(os.path.join(HERE, name), layers, photometric=2, orientation=o, rows_per_strip=3)
        out[name] = xref_model.load_tiff(os.path.join(HERE, name))
    np.savez_compressed(os.path.join(HERE, "tiff_orient_golden.npz"), **out)


def main():
    if "--only-orientation" in sys.argv:
        assert xref_model.available(), "build oracle/_ref first (make -C oracle ref)"
        orientation_fixtures()
        return
    assert xref.available() and xref_model.available(), "build oracle/_ref first (make -C oracle ref)"
    out = {}
    tmp = os.path.join(HERE, "_tmp.svo")
    for name, grid, cam, output, display, ratio, emission in render_cases():
        out[f"{name}/grid"] = grid
        out[f"{name}/camera"] = np.asarray(cam, dtype=np.float64)
        out[f"{name}/output"] = np.asarray(output)
        out[f"{name}/display"] = np.asarray(display)
        out[f"{name}/ratio"] = np.asarray(ratio, dtype=np.float64)
        out[f"{name}/emission"] = np.asarray(emission)
        # octrees come from the reference's own convert (sparse and rope)
        trees = {}
        for tname, ttype in (("sparse", 0), ("rope", 2)):
            xref_model.convert_mem(grid, tmp, chan_diff=0, type=ttype)
            raw = np.fromfile(tmp, dtype=np.uint8)
            trees[tname] = raw
            out[f"{name}/svo_{tname}"] = raw
        for t in TRAVERSALS:
            kw = dict(camera=cam, output=output, display=display, voxel_ratio=ratio, emission=emission)
            if t == "dda":
                img = xref.render("dda", grid=grid, **kw)
            else:
                raw = trees["rope" if t == "svo-rope" else "sparse"]
                side = int(np.frombuffer(raw[8:16].tobytes(), "<u8")[0])
                nodes = np.frombuffer(raw[24:].tobytes(), dtype=np.dtype([("c", "<u4", (8,)), ("col", "<u4"), ("d", "<u4")]))
                img = xref.render(t, nodes=nodes, side=side, **kw)
            out[f"{name}/image_{t}"] = img
    os.unlink(tmp)
    np.savez_compressed(os.path.join(HERE, "render_golden.npz"), **out)

    # convert: digests of the reference's .svo bytes for a grid of option combinations
    rng = np.random.default_rng(777)
    conv = {}
    grids = {"rand_20x9x5": random_grid(rng, 20, 9, 5, sparsity=0.3, quant=64), "blob_24": blobby_grid(rng, 24, 24, 24),
             "noise_8": rng.integers(0, 256, (8, 8, 8, 4), dtype=np.uint8)}
    for gname, g in grids.items():
        conv[f"{gname}/grid"] = g
        for ttype in (0, 1, 2):
            for hname, h in (("cd0", dict(chan_diff=0)), ("cd70", dict(chan_diff=70)), ("sd0", dict(std_dev=0.0)),
                             ("sd40", dict(std_dev=40.0))):
                st = xref_model.convert_mem(g, tmp, type=ttype, **h)
                raw = open(tmp, "rb").read()
                conv[f"{gname}/t{ttype}_{hname}/sha256"] = np.frombuffer(hashlib.sha256(raw).digest(), dtype=np.uint8)
                conv[f"{gname}/t{ttype}_{hname}/stats"] = np.asarray(
                    [st["total_leaves"], st["unique_leaves"], st["total_nodes"], st["depth"], len(raw)], dtype=np.uint64)
    os.unlink(tmp)
    np.savez_compressed(os.path.join(HERE, "convert_golden.npz"), **conv)

    # TIFF: files written by Pillow + what the real libtiff (through Grid::load_tiff) returns
    from PIL import Image
    rng = np.random.default_rng(5)
    vol = rng.integers(0, 256, (3, 5, 7, 4), dtype=np.uint8)
    tiff = {}
    for mode, arr in (("RGBA", vol), ("RGB", vol[..., :3]), ("L", vol[..., 0])):
        for big in (False, True):
            name = f"pillow_{mode}_{'big' if big else 'classic'}.tif"
            path = os.path.join(HERE, name)
            ims = [Image.fromarray(arr[z], mode) for z in range(arr.shape[0])]
            ims[0].save(path, save_all=True, append_images=ims[1:], big_tiff=big)
            tiff[name] = xref_model.load_tiff(path)
    # ... and hand-written files covering every sample layout (grey / grey+alpha / RGB / RGBA with
    # each ExtraSamples value): libtiff pre-multiplies only unassociated RGB, and ignores a grey
    # image's second sample unless ExtraSamples declares it alpha
    sys.path.insert(0, os.path.dirname(HERE))
    from util import TIFF_LAYOUTS, tiff_layout_name, write_tiff_layout
    for spp, phot, extra in TIFF_LAYOUTS:
        name = tiff_layout_name(spp, phot, extra)
        write_tiff_layout(os.path.join(HERE, name), spp, phot, extra)
        tiff[name] = xref_model.load_tiff(os.path.join(HERE, name))
    np.savez_compressed(os.path.join(HERE, "tiff_golden.npz"), **tiff)
    orientation_fixtures()
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
