"""Shared helpers for the test-suite: seeded volumes, cameras, comparison metrics."""
import struct

import numpy as np

CAM_SINGLE = ((0, 0, 1), (0, 1, 0), (0.5, 0.5, -1.5))  # reference camera-single.txt
CAM_ORBIT = ((0.125333, 0, 0.992115), (0, 1, 0), (0.249334, 0.5, -1.48423))  # camera.txt line 3
CAM_INSIDE = ((-0.0627904, 2.75301e-07, 0.998027), (-1.21213e-09, -1, 2.75769e-07),
              (0.531395, 0.5, 0.000985205))  # camera.txt line 100: origin inside the volume
CAM_OBLIQUE = ((0.3, -0.5, 0.7), (0.1, 1, 0.2), (0.2, 1.4, -0.3))
# looks at the centre of a volume stretched by --voxel-ratio 1:2:0.5 from outside, obliquely
CAM_ANISO = ((0.47, -0.6, 0.62), (0.1, 1, 0.2), (-0.6, 2.4, -1.2))
CAM_AXIS_NEG = ((0, 0, -1), (0, 1, 0), (0.5, 0.5, 2.5))
CAMERAS = {"single": CAM_SINGLE, "orbit": CAM_ORBIT, "inside": CAM_INSIDE, "oblique": CAM_OBLIQUE, "aniso": CAM_ANISO,
           "axis_neg": CAM_AXIS_NEG}


def random_grid(rng, nx, ny, nz, sparsity=0.5, quant=1):
    g = rng.integers(0, 256, (nz, ny, nx, 4), dtype=np.uint8)
    if quant > 1:
        g = (g // quant) * quant
    g[..., 3] = 255
    if sparsity > 0:
        g[rng.random((nz, ny, nx)) < sparsity] = (0, 0, 0, 255)
    return g


def blobby_grid(rng, nx, ny, nz):
    """Smooth-ish volume with large uniform regions (octrees with mixed leaf depths)."""
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    g = np.zeros((nz, ny, nx, 4), np.uint8)
    g[..., 3] = 255
    for _ in range(4):
        c = rng.random(3) * (nx, ny, nz)
        r = (0.15 + 0.25 * rng.random()) * min(nx, ny, nz)
        inside = (x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2 < r * r
        g[inside, :3] = rng.integers(32, 256, 3, dtype=np.uint8)
    return g


def image_diff(a, b):
    """(max abs channel difference, fraction of pixels with any channel differing by > 1)."""
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    return int(d.max()), float((d.max(axis=-1) > 1).mean())


def write_tiff(path, layers, *, photometric, extra=None, rows_per_strip=None, orientation=1, tile=None):
    """Minimal classic little-endian TIFF writer: layers = list of (H, W, spp) uint8 arrays."""
    out = bytearray(b"II*\x00\x00\x00\x00\x00")
    ifd_link = 4
    for a in layers:
        h, w, spp = a.shape
        chunks = []
        if tile:
            tw, th = tile
            for j in range(0, h, th):
                for i in range(0, w, tw):
                    t = np.zeros((th, tw, spp), np.uint8)
                    blk = a[j:j + th, i:i + tw]
                    t[:blk.shape[0], :blk.shape[1]] = blk
                    chunks.append(t.tobytes())
        else:
            rps = rows_per_strip or h
            for r in range(0, h, rps):
                chunks.append(a[r:r + rps].tobytes())
        offs = []
        for c in chunks:
            offs.append(len(out))
            out += c
            if len(out) & 1:
                out += b"\0"
        tags = [(256, 3, [w]), (257, 3, [h]), (258, 3, [8] * spp), (259, 3, [1]), (262, 3, [photometric]),
                (274, 3, [orientation]), (277, 3, [spp]), (284, 3, [1])]
        if tile:
            tags += [(322, 3, [tile[0]]), (323, 3, [tile[1]]), (324, 4, offs), (325, 4, [len(c) for c in chunks])]
        else:
            tags += [(273, 4, offs), (278, 3, [rows_per_strip or h]), (279, 4, [len(c) for c in chunks])]
        if extra is not None:
            tags.append((338, 3, [extra]))
        tags.sort()
        entries = bytearray()
        for tag, typ, vals in tags:
            fmt = {3: "H", 4: "I"}[typ]
            data = struct.pack("<%d%s" % (len(vals), fmt), *vals)
            if len(data) > 4:
                off = len(out)
                out += data
                if len(out) & 1:
                    out += b"\0"
                data = struct.pack("<I", off)
            entries += struct.pack("<HHI", tag, typ, len(vals)) + data.ljust(4, b"\0")
        ifd = len(out)
        out += struct.pack("<H", len(tags)) + entries + b"\0\0\0\0"
        out[ifd_link:ifd_link + 4] = struct.pack("<I", ifd)
        ifd_link = ifd + 2 + len(entries)
    with open(path, "wb") as f:
        f.write(out)


# (samples per pixel, photometric, ExtraSamples value or None): every chunky 8-bit layout the TIFF
# reader accepts; tests/golden/tiff_golden.npz holds what real libtiff returns for each
TIFF_LAYOUTS = [(1, 0, None), (1, 1, None), (2, 0, None), (2, 1, None), (2, 0, 0), (2, 1, 0), (2, 0, 1), (2, 1, 1),
                (2, 0, 2), (2, 1, 2), (3, 2, None), (4, 2, None), (4, 2, 0), (4, 2, 1), (4, 2, 2)]


def tiff_layout_name(spp, phot, extra):
    return f"layout_s{spp}_p{phot}_e{'x' if extra is None else extra}.tif"


def write_tiff_layout(path, spp, phot, extra):
    rng = np.random.default_rng(100 * spp + 10 * phot + (7 if extra is None else extra))
    layers = [rng.integers(0, 256, (4, 5, spp), dtype=np.uint8) for _ in range(2)]
    write_tiff(path, layers, photometric=phot, extra=extra, rows_per_strip=3)
