"""Tile slicing / normalisation (a11, a12) and the stitching back-end (a13-a17) through the C ABI."""
import numpy as np
import pytest

from oracle import cases, tiling_np as tl

PIPE_CASES, TILE_CASES = cases.PIPE_CASES, cases.TILE_CASES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def post():
    from yolo3_b200 import post_engine
    return post_engine(0)


@pytest.mark.parametrize("tag", list(TILE_CASES))
def test_tiles_normalized(post, golden, tag):
    from yolo3_b200 import tile_plan
    g = golden("tiling.npz")
    h, w, c, dt, tile, edge = TILE_CASES[tag]
    img = cases.synthetic_image(h, w, c, dt, seed=len(tag) * 7)
    xs, ys = tile_plan(h, w, tile, edge)
    assert xs.tolist() == g[tag + "_xs"].tolist() and ys.tolist() == g[tag + "_ys"].tolist()
    got = post.tiles_normalized(img, tile, edge)                        # [T, C, th, tw]
    tiles, _, _ = tl.cut_tiles(img, tile, edge)
    want = np.stack([tl.zscore(t.astype(np.float32)).transpose(2, 0, 1) for t in tiles])
    assert got.shape == want.shape
    # statistics are accumulated in fp64 on the GPU vs NumPy's fp32 pairwise sums: tolerance 2e-6 abs+rel
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=2e-6)
    probe = np.asarray([t[0, 5::97, 3::89].ravel()[:16] for t in got], np.float32)
    np.testing.assert_allclose(probe, g[tag + "_z_probe"], rtol=2e-6, atol=2e-6)


def test_flat_tile_branch(post):
    flat = np.full((64, 64, 1), 7, np.uint16)
    flat[0, 0, 0] = 8
    got = post.tiles_normalized(flat, (64, 64), 0)
    np.testing.assert_allclose(got[0, 0], tl.zscore(flat)[:, :, 0], rtol=0, atol=1e-6)


@pytest.mark.parametrize("tag", list(PIPE_CASES))
def test_stitch_pipeline_golden(post, golden, tag):
    """the reference's inference_image_tiled with an injected detector: bit-exact [n,6] float64"""
    g = golden("tiled_pipeline.npz")
    h, w, c, dt, tile, edge, nb, nc, minbox = PIPE_CASES[tag]
    img = cases.synthetic_image(h, w, c, dt, seed=100 + len(tag) + edge)
    fake = cases.FakeDetector(nb, nc, tile, seed=5 + edge)
    tiles, _, _ = tl.cut_tiles(img, tile, edge)
    dets = np.stack([fake(tl.zscore(t.astype(np.float32)).transpose(2, 0, 1)[None])[0] for t in tiles])
    pred = post.stitch_tiles(dets, (h, w), tile, minbox, edge)
    assert pred.dtype == np.float64 and np.array_equal(pred, g[tag + "_pred"])
    # sharded the way the multi-GPU path does it: two tile ranges, concatenated
    half = len(tiles) // 2
    if half:
        a = post.stitch_tiles(dets[:half], (h, w), tile, minbox, edge, first=0)
        b = post.stitch_tiles(dets[half:], (h, w), tile, minbox, edge, first=half)
        assert np.array_equal(np.concatenate([a, b]), g[tag + "_pred"])


@pytest.mark.parametrize("shape,dt", [((37, 53, 3), np.uint8), ((416, 416, 3), np.float32), ((1000, 1001), np.uint16),
                                       ((7,), np.int32), ((64, 64, 1), np.uint16)])
def test_zscore_normalize_any_shape(shape, dt):
    """imagereader.zscore_normalize (imagereader.py:34-46) has no multiple-of-32 rule: y3_zscore"""
    import imagereader
    rng = np.random.default_rng(len(shape) * 11 + shape[0])
    a = (rng.standard_normal(shape) * 40 + 100).astype(dt) if dt != np.uint16 else rng.integers(0, 65535, shape).astype(dt)
    got = imagereader.zscore_normalize(a)
    want = tl.zscore(a)
    assert got.dtype == np.float32 and got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=3e-6, atol=3e-6)
    # std <= 1 branch
    flat = np.full(shape, 5, dt)
    np.testing.assert_allclose(imagereader.zscore_normalize(flat), 0.0, atol=1e-6)
