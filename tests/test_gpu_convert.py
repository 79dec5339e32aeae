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
