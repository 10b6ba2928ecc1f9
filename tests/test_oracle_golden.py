"""The oracle (NumPy + C restatements) against the committed golden vectors that
oracle/gen_golden.py minted from the reference's own code.  CPU only."""
import hashlib

import numpy as np
import pytest

from oracle import cases, nms_c, postproc_np as pp, tiling_np as tl


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("tag", ["tiny", "small", "mid", "loose", "degen"])
@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_single_class_nms(golden, tag, impl):
    g = golden("nms_single.npz")
    fn = pp.greedy_nms if impl == "numpy" else nms_c.greedy_nms
    with np.errstate(all="ignore"):
        keep = fn(g[tag + "_boxes"], g[tag + "_scores"], float(g[tag + "_thr"]))
    assert keep == g[tag + "_keep"].tolist()


def test_iou_row(golden):
    g = golden("nms_single.npz")
    b = g["iou_boxes"]
    got = pp.iou_one_vs_many(b[0], b[1:])
    assert got.dtype == np.float32 and np.array_equal(got, g["iou_row0"])


@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_filter_and_per_class(golden, impl):
    g = golden("nms_per_class.npz")
    det = g["det"]
    f = pp.drop_small(det, 32)
    assert f.shape[0] == int(g["filtered_rows"]) and sha(f) == str(g["filtered_sha"])
    fn = pp.class_wise_nms if impl == "numpy" else nms_c.class_wise_nms
    b, s, l = fn(f[:, 0:4], f[:, 4:5], f[:, 5:])
    assert np.array_equal(b, g["pc_boxes"]) and np.array_equal(s, g["pc_scores"]) and np.array_equal(l, g["pc_labels"])
    assert l.dtype == np.int32 and s.dtype == np.float32
    b, s, l = fn(f[:, 0:4], f[:, 4:5], f[:, 5:], 0.45, 0.6)
    assert np.array_equal(b, g["pc45_boxes"]) and np.array_equal(s, g["pc45_scores"]) and np.array_equal(l, g["pc45_labels"])
    assert fn(f[:5, 0:4], f[:5, 4:5] * 0, f[:5, 5:]) == (None, None, None)


def test_multiclass_80(golden):
    g = golden("nms_per_class.npz")
    bm, om, cm = cases.multiclass_case(6000, 80, 900, seed=31)
    b, s, l = nms_c.class_wise_nms(bm, om, cm, 0.45, 0.1)
    assert b.shape[0] == int(g["mc80_k"])
    assert (sha(b), sha(s), sha(l)) == (str(g["mc80_sha_boxes"]), str(g["mc80_sha_scores"]), str(g["mc80_sha_labels"]))


TILE_CASES = cases.TILE_CASES


@pytest.mark.parametrize("tag", list(TILE_CASES))
def test_tiles_and_zscore(golden, tag):
    g = golden("tiling.npz")
    h, w, c, dt, tile, edge = TILE_CASES[tag]
    img = cases.synthetic_image(h, w, c, dt, seed=len(tag) * 7)
    tiles, xs, ys = tl.cut_tiles(img, tile, edge)
    assert xs == g[tag + "_xs"].tolist() and ys == g[tag + "_ys"].tolist()
    assert [sha(t) for t in tiles] == g[tag + "_tile_sha"].tolist()
    assert list(tiles[0].shape) == g[tag + "_tile_shape"].tolist()
    probe = np.asarray([tl.zscore(t.astype(np.float32))[5::97, 3::89, 0].ravel()[:16] for t in tiles], np.float32)
    assert np.array_equal(probe, g[tag + "_z_probe"])


def test_zscore_flat_branch(golden):
    g = golden("tiling.npz")
    flat = np.full((64, 64, 1), 7, np.uint16)
    flat[0, 0, 0] = 8
    assert np.array_equal(tl.zscore(flat)[:2, :2, 0], g["flat_z"])


PIPE_CASES = cases.PIPE_CASES


@pytest.mark.parametrize("tag", list(PIPE_CASES))
@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_tiled_pipeline(golden, tag, impl):
    g = golden("tiled_pipeline.npz")
    h, w, c, dt, tile, edge, nb, nc, minbox = PIPE_CASES[tag]
    img = cases.synthetic_image(h, w, c, dt, seed=100 + len(tag) + edge)
    fake = cases.FakeDetector(nb, nc, tile, seed=5 + edge)
    nms = pp.greedy_nms if impl == "numpy" else nms_c.greedy_nms
    pred = tl.tiled_inference(fake, img, tile, minbox, edge_range=edge, nms_fn=nms)
    assert pred.dtype == np.float64 and np.array_equal(pred, g[tag + "_pred"])


def test_k3_200k_c_oracle(golden):
    """Full-size K3 (200k boxes, 1 class, IoU 0.45): C oracle == reference verbatim (hash)."""
    g = golden("nms_k3.npz")
    b, s = cases.k3_single_class()
    assert sha(b) + sha(s) == str(g["in_sha"])
    keep = np.asarray(nms_c.greedy_nms(b, s, 0.45), dtype=np.int32)
    assert keep.size == int(g["n_keep"]) and sha(keep) == str(g["keep_sha"])
