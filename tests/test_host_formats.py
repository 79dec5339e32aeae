"""Host side of the path (no GPU): the C ABI loads and exports what include/xenodon_b200.h
declares; TIFF / SVO / convert / headless.conf / camera / stats / PNG behave like the
reference's code (checked against golden outputs of the reference and, where present, live
against oracle/_ref)."""
import hashlib
import os
import re
import struct
import zlib

import numpy as np
import pytest

from util import blobby_grid, random_grid

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_library_exports_every_declared_symbol(xb):
    header = open(os.path.join(ROOT, "include", "xenodon_b200.h")).read()
    declared = set(re.findall(r"XN_API\s+[\w\s\*]+?\b(xn_\w+)\s*\(", header))
    assert len(declared) >= 40
    import ctypes
    handle = ctypes.CDLL(xb.LIB_PATH)
    missing = [name for name in sorted(declared) if not hasattr(handle, name)]
    assert not missing, f"declared in the header but not exported: {missing}"
    assert declared == set(xb._PROTOTYPES), "the Python mirror binds exactly the header's entry points"
    assert b"sm_100a" in xb.lib().xn_version()


def test_no_gpu_means_loud_failure_not_fallback(xb):
    """Without a CUDA device every compute entry point fails with XN_ERR_CUDA."""
    try:
        n = xb.device_count()
    except xb.XenodonError as e:
        assert e.status == -2
        n = 0
    if n == 0:
        with pytest.raises(xb.XenodonError):
            xb.Context(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "xenodon_b200")
    for base, _, files in os.walk(pkg):
        if os.path.basename(base) in ("build", "__pycache__"):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(base, f), errors="replace").read()
                assert "xn_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f


def test_unorm8_reciprocal_identity():
    """csrc/xn_device.cuh unorm8(): q = b*r, rem = fma(-q, 255, b), q + rem*r == b / 255 exactly."""
    b = np.arange(256, dtype=np.float32)
    r = np.float32(1.0) / np.float32(255.0)
    q = (b * r).astype(np.float64)
    rem = (b.astype(np.float64) - q * 255.0).astype(np.float32).astype(np.float64)  # fma: one rounding
    out = (q + rem * np.float64(r)).astype(np.float32)
    assert np.array_equal(out, b / np.float32(255.0))


# ---- TIFF ----
def test_tiff_reader_matches_real_libtiff_golden(xb):
    z = np.load(os.path.join(GOLD, "tiff_golden.npz"))
    assert len(z.files) == 6 + 15  # Pillow-written files + one hand-written file per sample layout
    for name in z.files:
        mine = xb.Grid.load_tiff(os.path.join(GOLD, name)).data
        assert np.array_equal(mine, z[name]), name  # bottom-up rows, premultiplied unassociated alpha


def test_tiff_orientation_tag_matches_real_libtiff(xb):
    """All eight values of the Orientation tag: libtiff's TIFFReadRGBAImage reverses rows for 1 / 5,
    rows and columns for 2 / 6, columns for 3 / 7 and nothing for 4 / 8 (it never transposes);
    fixtures written by tests/golden/make_golden.py from the real library."""
    z = np.load(os.path.join(GOLD, "tiff_orient_golden.npz"))
    assert len(z.files) == 8
    seen = set()
    for name in z.files:
        got = xb.Grid.load_tiff(os.path.join(GOLD, name)).data
        assert np.array_equal(got, z[name]), name
        seen.add(got.tobytes())
        info = xb.tiff_stream_info(os.path.join(GOLD, name))
        o = int(name.split("_")[1].split(".")[0])
        assert info["flip"] == (o in (1, 5, 2, 6)) and info["mirror"] == (o in (2, 6, 3, 7)), name
    assert len(seen) == 4


def test_tiff_with_crafted_dimensions_is_rejected_not_wrapped(xb, tmp_path):
    """Width * height * 4 * layers must not wrap: a directory claiming 2^31 x 2^31 pixels (its strips
    all pointing at the same few bytes) is refused with XN_ERR_LIMIT before any size is computed from it."""
    def ifd(w, h):
        out = bytearray(b"II*\x00\x08\x00\x00\x00")
        tags = [(256, 4, w), (257, 4, h), (258, 3, 8), (259, 3, 1), (262, 3, 1), (273, 4, 8), (277, 3, 1),
                (278, 4, h), (279, 4, 16)]
        out += struct.pack("<H", len(tags))
        for t, ty, v in tags:
            out += struct.pack("<HHI", t, ty, 1) + (struct.pack("<I", v) if ty == 4 else struct.pack("<HH", v, 0))
        out += struct.pack("<I", 0)
        return bytes(out)
    for w, h in ((0x80000000, 0x80000000), (0xFFFFFFFF, 3), (3, 0x80000000), (0x7FFFFFFF, 0x7FFFFFFF)):
        p = tmp_path / f"crafted_{w}_{h}.tif"
        p.write_bytes(ifd(w, h))
        for call in (lambda: xb.Grid.load_tiff(p), lambda: xb.tiff_stream_info(p)):
            with pytest.raises(xb.XenodonError) as e:
                call()
            assert e.value.status == -5, (w, h, str(e.value))  # XN_ERR_LIMIT


def test_tiff_write_read_roundtrip_and_errors(xb, tmp_path):
    rng = np.random.default_rng(1)
    g = rng.integers(0, 256, (4, 6, 5, 4), dtype=np.uint8)
    for big in (False, True):
        p = tmp_path / f"v{int(big)}.tif"
        xb.Grid(g).save_tiff(p, bigtiff=big)
        back = xb.Grid.load_tiff(p)
        assert back.dimensions == (5, 6, 4) and np.array_equal(back.data, g)
    with pytest.raises(xb.XenodonError, match="Failed to open"):
        xb.Grid.load_tiff(tmp_path / "missing.tif")
    bad = tmp_path / "bad.tif"
    bad.write_bytes(b"not a tiff at all")
    with pytest.raises(xb.XenodonError, match="Failed to open"):
        xb.Grid.load_tiff(bad)
    trunc = tmp_path / "trunc.tif"
    trunc.write_bytes((tmp_path / "v0.tif").read_bytes()[:200])
    with pytest.raises(xb.XenodonError):
        xb.Grid.load_tiff(trunc)


def test_tiff_reader_live_against_reference_loader(xb, tmp_path):
    from oracle import xref_model
    if not xref_model.available():
        pytest.skip("oracle/_ref not built on this machine; tiff_golden.npz covers it")
    from PIL import Image
    rng = np.random.default_rng(2)
    vol = rng.integers(0, 256, (4, 9, 11, 4), dtype=np.uint8)
    for mode, arr in (("RGBA", vol), ("RGB", vol[..., :3]), ("L", vol[..., 0])):
        p = tmp_path / f"{mode}.tif"
        ims = [Image.fromarray(arr[z], mode) for z in range(arr.shape[0])]
        ims[0].save(p, save_all=True, append_images=ims[1:], big_tiff=True)
        assert np.array_equal(xb.Grid.load_tiff(p).data, xref_model.load_tiff(p))
    p = tmp_path / "ours.tif"
    xb.Grid(vol).save_tiff(p)
    assert np.array_equal(xref_model.load_tiff(p), vol)  # the reference reads what we write


# ---- SVO ----
def test_svo_roundtrip_and_validation(xb, tmp_path):
    rng = np.random.default_rng(3)
    tree, _ = xb.build_octree(xb.Grid(random_grid(rng, 8, 8, 8)), chan_diff=0)
    p = tmp_path / "t.svo"
    tree.save_svo(p)
    raw = p.read_bytes()
    assert raw[:8] == b"XNDN-SVO" and struct.unpack("<QQ", raw[8:24]) == (8, len(tree.nodes))
    assert len(raw) == 24 + 40 * len(tree.nodes)
    back = xb.Octree.load_svo(p)
    assert back.side == 8 and np.array_equal(back.nodes, tree.nodes)
    (tmp_path / "short.svo").write_bytes(raw[:-4])
    with pytest.raises(xb.XenodonError, match="File size does not match number of nodes"):
        xb.Octree.load_svo(tmp_path / "short.svo")
    (tmp_path / "magic.svo").write_bytes(b"XNDN-SVX" + raw[8:])
    with pytest.raises(xb.XenodonError, match="Invalid format id"):
        xb.Octree.load_svo(tmp_path / "magic.svo")
    with pytest.raises(xb.XenodonError, match="Failed to open"):
        xb.Octree.load_svo(tmp_path / "nope.svo")
    from oracle import xref_model
    if xref_model.available():  # the reference's own reader accepts our file and rewrites it identically
        side, count = xref_model.svo_roundtrip(p, tmp_path / "ref.svo")
        assert (side, count) == (8, len(tree.nodes)) and (tmp_path / "ref.svo").read_bytes() == raw


def test_convert_matches_reference_digests(xb):
    z = np.load(os.path.join(GOLD, "convert_golden.npz"))
    checked = 0
    for gname in sorted({k.split("/")[0] for k in z.files}):
        g = xb.Grid(z[f"{gname}/grid"])
        for ttype in (0, 1, 2):
            for hname, h in (("cd0", dict(chan_diff=0)), ("cd70", dict(chan_diff=70)), ("sd0", dict(std_dev=0.0)),
                             ("sd40", dict(std_dev=40.0))):
                tree, st = xb.build_octree(g, type=ttype, **h)
                raw = b"XNDN-SVO" + struct.pack("<QQ", tree.side, len(tree.nodes)) + tree.nodes.tobytes()
                ref = [int(v) for v in z[f"{gname}/t{ttype}_{hname}/stats"]]
                assert [st["total_leaves"], st["unique_leaves"], st["total_nodes"], st["depth"], len(raw)] == ref
                assert hashlib.sha256(raw).digest() == z[f"{gname}/t{ttype}_{hname}/sha256"].tobytes()
                checked += 1
    assert checked == 36


@pytest.mark.parametrize("dims", [(1, 1, 1), (3, 3, 3), (16, 16, 16), (40, 33, 17), (64, 64, 64)])
def test_convert_matches_oracle_builder(xb, xo, dims):
    """Bottom-up O(N) construction == the reference's top-down rescanning construction, byte for byte."""
    rng = np.random.default_rng(sum(dims))
    for g in (random_grid(rng, *dims, quant=64), blobby_grid(rng, *dims)):
        for ttype in (xb.TYPE_SPARSE, xb.TYPE_DAG, xb.TYPE_ROPE):
            for h in (dict(chan_diff=0), dict(chan_diff=90), dict(std_dev=25.0)):
                tree, st = xb.build_octree(xb.Grid(g), type=ttype, **h)
                nodes, side, ost = xo.build_octree(g, type=ttype, **h)
                assert side == tree.side and st == ost
                assert tree.nodes.tobytes() == nodes.tobytes()


def test_convert_is_deterministic_and_validates(xb):
    rng = np.random.default_rng(8)
    g = xb.Grid(blobby_grid(rng, 24, 24, 24))
    a, _ = xb.build_octree(g, chan_diff=0, type=xb.TYPE_DAG)
    b, _ = xb.build_octree(g, chan_diff=0, type=xb.TYPE_DAG)
    assert a.nodes.tobytes() == b.nodes.tobytes()
    sparse, _ = xb.build_octree(g, chan_diff=0)
    assert len(a.nodes) < len(sparse.nodes)  # the DAG merges identical subtrees
    with pytest.raises(ValueError):
        xb.build_octree(g, chan_diff=0, std_dev=1.0)


# ---- headless.conf ----
def test_headless_config_grammar(xb):
    text = ("device { vkindex = 0 offset = (0, 0) extent = (960, 1080) }\n"
            "device {\n    extent = (960,1080)\n    vkindex = 1\n    offset = (960,0)\n}\n")
    assert xb.parse_headless_config(text) == [(0, (0, 0, 960, 1080)), (1, (960, 0, 960, 1080))]
    assert xb.rect_union([r for _, r in xb.parse_headless_config(text)]) == (0, 0, 1920, 1080)
    cases = {
        "": "At least one device entry is required",
        "device { vkindex = 0 offset = (0,0) }": "Missing key 'extent'",
        "device { vkindex = 0 vkindex = 1 offset = (0,0) extent = (1,1) }": "Ambiguous key 'vkindex'",
        "# comment\ndevice { vkindex = 0 offset = (0,0) extent = (1,1) }": "Unexpected character '#'",
        "device { vkindex = 0 offset = (0,0) extent = (1,1) } trailing": "Unexpected key 'trailing'",
        "gpu { vkindex = 0 }": "Unexpected key 'gpu'",
        "device { vkindex = x offset = (0,0) extent = (1,1) }": "Expected numeric character",
        "device { vkindex = 0 offset = (0;0) extent = (1,1) }": "Expected character ','",
        "device { vkindex = 0 offset = (0,0) extent = (1,1)": "Unexpected end of input",
    }
    for bad, msg in cases.items():
        with pytest.raises(xb.XenodonError, match=re.escape(msg)):
            xb.parse_headless_config(bad)


# ---- camera scripts ----
def test_camera_script_parser(xb):
    one = xb.parse_camera_script("0 0 1 0 1 0 0.5 0.5 -1.5")  # no trailing newline, like camera-single.txt
    assert one.shape == (1, 3, 3) and tuple(one[0, 2]) == (0.5, 0.5, -1.5)
    two = xb.parse_camera_script("0 0 1 0 1 0 0.5 0.5 -1.5\n1e-3 0 1\n0 1 0\n0.5 0.5 2.5\n\n")
    assert two.shape == (2, 3, 3) and abs(two[1, 0, 0] - 1e-3) < 1e-9
    for bad in ("", "0 0 1 0 1 0 0.5 0.5", "0 0 1 0 1 0 0.5 0.5 x"):
        with pytest.raises(xb.XenodonError, match="Syntax error in camera input file"):
            xb.parse_camera_script(bad)


def test_generated_camera_scripts_follow_the_reference_files(xb):
    from xenodon_b200 import cameras
    assert cameras.camera_single().shape == (1, 3, 3)
    assert cameras.camera_rotate().shape == (150, 3, 3) and cameras.camera_benchmark().shape == (150, 3, 3)
    # re-parsing the generated text gives the generated frames
    for name, fn in cameras.SCRIPTS.items():
        frames = fn()
        assert np.allclose(xb.parse_camera_script(cameras.to_text(frames)), frames, atol=1e-6)
    # frames 100-149 of the benchmark path have the origin 0.5 from the centre (inside the volume)
    pos = cameras.camera_benchmark()[100:, 2]
    assert np.allclose(np.linalg.norm(pos - 0.5, axis=1), 0.5, atol=1e-4)
    ref_dir = "/root/reference"
    if os.path.isdir(ref_dir):
        for name, fn in cameras.SCRIPTS.items():
            ref = xb.parse_camera_script(open(os.path.join(ref_dir, name + ".txt")).read())
            assert ref.shape == fn().shape and np.abs(ref - fn()).max() < 2e-5


# ---- stats + PNG ----
def test_stats_file_layout(xb, tmp_path):
    frames = [xb.RenderStats(2073600, 1, 2.5, 2.5, 2.5), xb.RenderStats(2073600, 2, 4.0, 3.0, 1.0)]
    p = tmp_path / "stats.txt"
    xb.write_stats(p, frames, 0.5)
    lines = p.read_text().splitlines()
    assert lines[0] == "total rays: 4147200"
    assert lines[1] == "total render time: 6.5"
    assert lines[2].startswith("total mray/s: ") and abs(float(lines[2].split(": ")[1]) - 4147200 / 6500) < 1e-9
    assert lines[3] == "average fps: 4" and lines[4] == "frames: 2"
    assert lines[5].startswith("# Frame number: total rays, outputs, total render time")
    assert lines[6] == "frame 0: 2073600 rays, 1, 2.5 ms, 2.5 ms, 2.5 ms, 829.44 mray/s"
    assert lines[7] == "frame 1: 2073600 rays, 2, 4 ms, 3 ms, 1 ms, 518.4 mray/s"


def test_png_writer(xb, tmp_path):
    from PIL import Image
    rng = np.random.default_rng(4)
    img = rng.integers(0, 256, (17, 23, 4), dtype=np.uint8)
    p = tmp_path / "f.png"
    xb.write_png(p, img)
    raw = p.read_bytes()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n" and zlib.crc32(raw[12:29]) == struct.unpack(">I", raw[29:33])[0]
    assert np.array_equal(np.asarray(Image.open(p)), img)


def test_synthetic_volumes_have_the_named_shapes(xb):
    b = xb.Grid.synthetic(xb.SYNTH_BUNNY, 64, 45, 64).data
    assert (b[..., 3] == 255).all() and (b[..., 0] == b[..., 1]).all() and (b[..., 1] == b[..., 2]).all()
    frac = (b[..., 0] > 0).mean()
    assert 0.10 < frac < 0.35 and b[..., 0][b[..., 0] > 0].min() > 5  # CT threshold: values <= 5 are 0
    corner = b[:4, :, :4, 0]
    assert not corner.any()  # outside the cylinder mask
    t = xb.Grid.synthetic(xb.SYNTH_TNG, 64, 64, 64).data
    floor = (t[..., :3] == (0, 0, 3)).all(axis=-1).mean()
    assert 0.8 < floor < 0.97 and (t[..., 3] == 255).all()
    assert np.array_equal(t, xb.Grid.synthetic(xb.SYNTH_TNG, 64, 64, 64).data)  # deterministic
    assert not np.array_equal(t, xb.Grid.synthetic(xb.SYNTH_TNG, 64, 64, 64, seed=7).data)


@pytest.mark.parametrize("dims", [(16, 16, 16), (33, 20, 9), (5, 70, 3), (40, 8, 129), (512, 361, 512)])
def test_bricked_layout_index_function(xb, dims):
    """csrc/xn_brick.h: the slot index is a bijection onto [0, total), a 32-byte sector holds a
    2x2x2 voxel cube, the axis padded worst sits on top (no power-of-two padding there), and a DDA
    step is (d + K) & mask on the axis' own bits -- including the wrap at -1 and n."""
    nx, ny, nz = dims
    L = xb.brick_layout(nx, ny, nz)
    mx, my, mz = L["mask"]
    assert mx & my == 0 and mx & mz == 0 and my & mz == 0
    nb = [(n + 7) // 8 for n in dims]
    pad = [(1 << max(0, (b - 1).bit_length())) / b for b in nb]
    assert pad[L["top"]] == max(pad)
    lower = [a for a in range(3) if a != L["top"]]
    total = 512 * nb[L["top"]]
    for a in lower:
        total *= 1 << max(0, (nb[a] - 1).bit_length())
    assert L["total"] == total
    if nx * ny * nz <= 1 << 18:
        z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
        xyz = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1)
    else:
        xyz = np.random.default_rng(1).integers(0, dims, (200000, 3))
        xyz = np.unique(xyz, axis=0)
    idx = xb.brick_indices(nx, ny, nz, xyz)
    assert idx.max() < L["total"] and len(np.unique(idx)) == len(idx)
    # one sector (8 consecutive slots) = one 2x2x2 cube; one 2 KiB brick = one 8x8x8 cube
    assert np.array_equal(idx >> np.uint64(3) == (idx[0] >> np.uint64(3)), (xyz >> 1 == xyz[0] >> 1).all(axis=1))
    assert np.array_equal(idx >> np.uint64(9) == (idx[0] >> np.uint64(9)), (xyz >> 3 == xyz[0] >> 3).all(axis=1))
    # stepping: x from -1 to nx (the kernel's cursor passes through both while out of range)
    for axis, n in enumerate(dims):
        mask = L["mask"][axis] & 0xFFFFFFFFFFFFFFFF
        pts = np.zeros((n + 2, 3), np.int32)
        pts[:, axis] = np.arange(-1, n + 1)
        d = [int(v) & mask for v in xb.brick_indices(nx, ny, nz, pts)]
        for i in range(len(d) - 1):
            assert (d[i] - mask) & 0xFFFFFFFFFFFFFFFF & mask == d[i + 1]      # +1: K = -mask
            assert (d[i + 1] - 1) & 0xFFFFFFFFFFFFFFFF & mask == d[i]          # -1: K = -1
    # a forced top axis (the 64-bit cursor wants z on top) is honoured
    assert xb.brick_layout(nx, ny, nz, 2)["top"] == 2


def test_ingest_plan_of_tiff_layouts(xb, tmp_path):
    """What the streaming ingest (xn_upload_grid_tiff) decides per file, checked without a GPU: strip
    files of one sample layout stream (adjacent strips merge into one read per layer), tiled or
    mixed-layout files fall back to the host decoder, and the decode flags follow libtiff's rules."""
    from util import write_tiff
    rng = np.random.default_rng(8)
    layers = [rng.integers(0, 256, (12, 10, 4), dtype=np.uint8) for _ in range(3)]
    p = tmp_path / "strips.tif"
    write_tiff(p, layers, photometric=2, extra=2, rows_per_strip=5)
    info = xb.tiff_stream_info(p)
    assert info == {"streamable": True, "samples": 4, "photometric": 2, "has_alpha": True, "unassociated": True,
                    "flip": True, "mirror": False, "runs": 3}
    write_tiff(p, layers, photometric=2, extra=1, orientation=4)
    info = xb.tiff_stream_info(p)
    assert info["streamable"] and not info["unassociated"] and not info["flip"]
    grey = [a[..., :2].copy() for a in layers]
    write_tiff(p, grey, photometric=1, extra=None)  # second sample not declared alpha: ignored
    info = xb.tiff_stream_info(p)
    assert info["streamable"] and info["samples"] == 2 and not info["has_alpha"]
    write_tiff(p, grey, photometric=0, extra=2)      # grey + unassociated alpha: alpha kept, no pre-multiplication
    info = xb.tiff_stream_info(p)
    assert info["has_alpha"] and not info["unassociated"] and info["photometric"] == 0
    write_tiff(p, layers, photometric=2, extra=2, tile=(16, 16))
    assert not xb.tiff_stream_info(p)["streamable"]
    # our own writer's files stream, one run per layer
    g = rng.integers(0, 256, (4, 6, 5, 4), dtype=np.uint8)
    xb.Grid(g).save_tiff(p)
    info = xb.tiff_stream_info(p)
    assert info["streamable"] and info["runs"] == 4 and info["samples"] == 4
    with pytest.raises(xb.XenodonError, match="Failed to open"):
        xb.tiff_stream_info(tmp_path / "missing.tif")
