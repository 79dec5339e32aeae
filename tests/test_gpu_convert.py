"""`xenodon convert` on the GPU (xn_convert_resident_grid): the node array must be byte-identical
to the oracle's restatement of the reference builder (oracle/xn_oracle.c, itself pinned against the
reference's own src/model code) and to the product's host builder, for sparse and rope trees, any
threshold, grids that are not powers of two, and through to rendering."""
import numpy as np
import pytest

from util import CAMERAS, blobby_grid, random_grid

pytestmark = pytest.mark.gpu


def _gpu_convert(xb, g, **kw):
    ctx = xb.Context(0)
    try:
        ctx.upload_grid(xb.Grid(g))
        return ctx.convert_resident_grid(want_nodes=True, bind=False, **kw)
    finally:
        ctx.close()


@pytest.mark.parametrize("dims", [(1, 1, 1), (2, 2, 2), (3, 3, 3), (8, 8, 8), (16, 11, 16), (20, 9, 5), (40, 33, 17),
                                  (64, 64, 64), (33, 70, 12)])
def test_gpu_convert_matches_host_builder(xb, xo, dims):
    rng = np.random.default_rng(sum(dims) * 7 + 1)
    grids = [random_grid(rng, *dims, quant=64), blobby_grid(rng, *dims), np.full((dims[2], dims[1], dims[0], 4), 77, np.uint8),
             rng.integers(0, 256, (dims[2], dims[1], dims[0], 4), dtype=np.uint8)]
    for g in grids:
        for ttype in (xb.TYPE_SPARSE, xb.TYPE_DAG, xb.TYPE_ROPE):
            for thr in (0, 60, 255):
                tree, st, count, side = _gpu_convert(xb, g, chan_diff=thr, type=ttype)
                onodes, oside, ost = xo.build_octree(g, chan_diff=thr, type=ttype)  # the oracle, directly
                assert (count, side) == (len(onodes), oside)
                assert tree.nodes.tobytes() == onodes.tobytes(), (dims, ttype, thr)
                assert st == ost, (dims, ttype, thr, st, ost)
                ref, rst = xb.build_octree(xb.Grid(g), chan_diff=thr, type=ttype)  # and the product's host builder
                assert tree.nodes.tobytes() == ref.nodes.tobytes() and st == rst, (dims, ttype, thr)


def test_gpu_convert_synthetic_volumes_and_render(xb, xo):
    """Device-generated volume -> GPU convert -> bound octree -> every SVO traversal == oracle on the
    host-built tree of the host-generated (bit-identical) volume."""
    for kind, dims in ((xb.SYNTH_BUNNY, (96, 68, 96)), (xb.SYNTH_TNG, (128, 128, 128))):
        host = xb.Grid.synthetic(kind, *dims)
        ref, rst = xb.build_octree(host, chan_diff=0, type=xb.TYPE_ROPE)
        ctx = xb.Context(0)
        ctx.synth_grid(kind, *dims)
        tree, st, count, side = ctx.convert_resident_grid(chan_diff=0, type=xb.TYPE_ROPE, bind=True, want_nodes=True)
        assert tree.nodes.tobytes() == ref.nodes.tobytes() and st == rst
        ctx.set_precision(True)
        ctx.set_target((0, 0, 160, 90))
        ctx.set_params((1, 1, 1), None, 4.0)
        for t in ("svo-naive", "svo-df", "esvo", "svo-rope"):
            ctx.render(t, CAMERAS["orbit"])
            ctx.sync()
            img = ctx.download()
            want = xo.render(t, nodes=ref.nodes, side=ref.side, camera=CAMERAS["orbit"], output=(0, 0, 160, 90),
                             emission=4.0, want_stats=False)[0]
            assert np.array_equal(img, want), t
        ctx.close()


def test_gpu_dag_matches_reference_digests_and_renders(xb, xo):
    """`convert --dag` on the GPU: the SHA-256 of the .svo bytes equals the digest of the reference's
    own output (tests/golden/convert_golden.npz, written from its src/model code), for every
    --chan-diff fixture; the bunny-shape volume's DAG (8x fewer nodes than the sparse tree) is
    byte-identical to the oracle's and renders bit-identically through every DAG-capable traversal."""
    import hashlib
    import os
    import struct
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "convert_golden.npz"))
    checked = 0
    for gname in ("rand_20x9x5", "blob_24", "noise_8"):
        g = z[f"{gname}/grid"]
        for hname, thr in (("cd0", 0), ("cd70", 70)):
            tree, st, count, side = _gpu_convert(xb, g, chan_diff=thr, type=xb.TYPE_DAG)
            raw = b"XNDN-SVO" + struct.pack("<QQ", side, count) + tree.nodes.tobytes()
            assert hashlib.sha256(raw).digest() == z[f"{gname}/t1_{hname}/sha256"].tobytes(), (gname, hname)
            want = z[f"{gname}/t1_{hname}/stats"]
            assert [st["total_leaves"], st["unique_leaves"], st["total_nodes"], st["depth"]] == list(want[:4]), (gname, hname)
            checked += 1
    assert checked == 6
    host = xb.Grid.synthetic(xb.SYNTH_BUNNY, 96, 68, 96)
    onodes, oside, ost = xo.build_octree(host.data, chan_diff=0, type=xo.TYPE_DAG)
    ctx = xb.Context(0)
    try:
        ctx.synth_grid(xb.SYNTH_BUNNY, 96, 68, 96)
        tree, st, count, side = ctx.convert_resident_grid(chan_diff=0, type=xb.TYPE_DAG, bind=True, want_nodes=True)
        assert tree.nodes.tobytes() == onodes.tobytes() and st == ost and count < st["total_nodes"]
        ctx.set_precision(True)
        ctx.set_target((0, 0, 160, 90))
        ctx.set_params((1, 1, 1), None, 4.0)
        for t in ("svo-naive", "svo-df", "esvo"):
            ctx.render(t, CAMERAS["orbit"])
            ctx.sync()
            want_img = xo.render(t, nodes=onodes, side=oside, camera=CAMERAS["orbit"], output=(0, 0, 160, 90), emission=4.0,
                                 want_stats=False)[0]
            assert np.array_equal(ctx.download(), want_img), t
    finally:
        ctx.close()


def test_gpu_std_dev_matches_reference_digests_and_oracle(xb, xo):
    """`convert --std-dev` on the GPU decides every split from exact integer sums plus a rigorous
    bound on the rounding of the reference's binary64 evaluation: the .svo bytes equal the
    reference's digests (sd0, sd40 fixtures, all three tree types) and the oracle's on random /
    smooth volumes for a sweep of thresholds; a threshold that coincides with a cell's deviation
    (0.5 on a cell of four 0s and four 1s) is refused with XN_ERR_LIMIT, never guessed."""
    import hashlib
    import os
    import struct
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "convert_golden.npz"))
    for gname in ("rand_20x9x5", "blob_24", "noise_8"):
        g = z[f"{gname}/grid"]
        for ttype in (0, 1, 2):
            for hname, sd in (("sd0", 0.0), ("sd40", 40.0)):
                tree, st, count, side = _gpu_convert(xb, g, std_dev=sd, type=ttype)
                raw = b"XNDN-SVO" + struct.pack("<QQ", side, count) + tree.nodes.tobytes()
                assert hashlib.sha256(raw).digest() == z[f"{gname}/t{ttype}_{hname}/sha256"].tobytes(), (gname, ttype, hname)
    rng = np.random.default_rng(99)
    for g in (random_grid(rng, 33, 20, 17, quant=16), blobby_grid(rng, 40, 29, 33),
              rng.integers(0, 256, (16, 16, 16, 4), dtype=np.uint8)):
        for sd in (0.0, 0.3, 7.77, 31.0, 90.0, 300.0):
            tree, st, count, side = _gpu_convert(xb, g, std_dev=sd, type=xb.TYPE_SPARSE)
            onodes, oside, ost = xo.build_octree(g, std_dev=sd, type=xo.TYPE_SPARSE)
            assert tree.nodes.tobytes() == onodes.tobytes() and st == ost, sd
    g = np.zeros((2, 2, 2, 4), np.uint8)
    g[0, :, :, 0] = 1  # channel r: four 0s, four 1s -> deviation exactly 0.5
    with pytest.raises(xb.XenodonError) as e:
        _gpu_convert(xb, g, std_dev=0.5)
    assert e.value.status == -5 and "host builder" in str(e.value)
    tree, _, count, _ = _gpu_convert(xb, g, std_dev=0.4999)
    assert count == 9
    tree, _, count, _ = _gpu_convert(xb, g, std_dev=0.5001)
    assert count == 1


def test_gpu_convert_rejects_bad_arguments(xb):
    ctx = xb.Context(0)
    ctx.upload_grid(xb.Grid(np.zeros((4, 4, 4, 4), np.uint8)))
    with pytest.raises(xb.XenodonError, match="unknown octree type"):
        ctx.convert_resident_grid(type=7)
    ctx.close()
    ctx = xb.Context(0)
    with pytest.raises(xb.XenodonError, match="no grid is resident"):
        ctx.convert_resident_grid()
    ctx.close()
