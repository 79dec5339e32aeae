"""Shared helpers for the test-suite: seeded volumes, cameras, comparison metrics."""
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
