"""CUDA NMS family through the C ABI vs the golden vectors (reference verbatim) and the C oracle.
Kept-index sets must be BIT-EXACT (north_star)."""
import hashlib

import numpy as np
import pytest

from oracle import cases, nms_c, postproc_np as pp

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def post():
    from yolo3_b200 import post_engine
    return post_engine(0)


@pytest.mark.parametrize("tag", ["tiny", "small", "mid", "loose", "degen"])
def test_single_class_golden(post, golden, tag):
    g = golden("nms_single.npz")
    keep = post.single_class_nms(g[tag + "_boxes"], g[tag + "_scores"], float(g[tag + "_thr"]))
    assert keep.tolist() == g[tag + "_keep"].tolist()


def test_iou_row_golden(post, golden):
    g = golden("nms_single.npz")
    b = g["iou_boxes"]
    assert np.array_equal(post.compute_iou(b[0], b[1:]), g["iou_row0"])


def test_filter_and_per_class_golden(post, golden):
    g = golden("nms_per_class.npz")
    f = post.filter_small(g["det"], 32)
    assert f.shape[0] == int(g["filtered_rows"]) and sha(f) == str(g["filtered_sha"])
    b, s, l = post.per_class_nms(f[:, 0:4], f[:, 4:5], f[:, 5:])
    assert np.array_equal(b, g["pc_boxes"]) and np.array_equal(s, g["pc_scores"]) and np.array_equal(l, g["pc_labels"])
    b, s, l = post.per_class_nms(f[:, 0:4], f[:, 4:5], f[:, 5:], 0.45, 0.6)
    assert np.array_equal(b, g["pc45_boxes"]) and np.array_equal(s, g["pc45_scores"]) and np.array_equal(l, g["pc45_labels"])
    assert post.per_class_nms(f[:5, 0:4], f[:5, 4:5] * 0, f[:5, 5:]) == (None, None, None)


def test_multiclass_80_golden(post, golden):
    g = golden("nms_per_class.npz")
    bm, om, cm = cases.multiclass_case(6000, 80, 900, seed=31)
    b, s, l = post.per_class_nms(bm, om, cm, 0.45, 0.1)
    assert b.shape[0] == int(g["mc80_k"])
    assert (sha(b), sha(s), sha(l)) == (str(g["mc80_sha_boxes"]), str(g["mc80_sha_scores"]), str(g["mc80_sha_labels"]))


@pytest.mark.parametrize("m", [1, 2, 31, 32, 33, 64, 65, 100, 128, 129, 200, 256, 257, 511, 512, 513, 1025, 4096, 5632, 5633, 8192, 8193,
                               20000, 24576, 24577, 40000])
def test_sizes_vs_c_oracle(post, m):
    """every route of the segmented pipeline: one warp per segment (<= 256 boxes, 1 / 2 / 4 / 8 boxes per lane), one CTA with
    the small (<= 5632) and the large (<= 24576) key buffer, chunk boundaries of the bitmask NMS (512), and the global-sort
    pipeline above that"""
    b, s = cases.nms_case(m, 40 * np.sqrt(m) + 50, seed=m, wh=(20, 120))
    for thr in (0.3, 0.45):
        assert post.single_class_nms(b, s, thr).tolist() == nms_c.greedy_nms(b, s, thr)


def test_empty(post):
    assert post.single_class_nms(np.zeros((0, 4), np.float32), np.zeros(0, np.float32), 0.3).size == 0


def test_ties_follow_documented_rule(post):
    """duplicate scores: score desc, then index asc - same rule as the oracle"""
    rng = np.random.default_rng(5)
    b = cases.boxes_on_canvas(3000, 400, 20, 90, rng)
    s = rng.integers(1, 40, 3000).astype(np.float32) / 40
    assert post.single_class_nms(b, s, 0.3).tolist() == nms_c.greedy_nms(b, s, 0.3)


def test_heavy_multiclass_vs_c_oracle(post):
    b, o, c = cases.multiclass_case(3000, 10, 500, seed=77, dominant_only=False)
    R = nms_c.class_wise_nms(b, o, c, 0.45, 0.1)
    G = post.per_class_nms(b, o, c, 0.45, 0.1)
    assert all(np.array_equal(x, y) for x, y in zip(R, G))


def test_k3_200k_single_class(post, golden):
    """BASELINE config 3: 200k candidates, 1 class, IoU 0.45 - bit-exact vs the reference's own run"""
    g = golden("nms_k3.npz")
    b, s = cases.k3_single_class()
    keep = post.single_class_nms(b, s, 0.45)
    assert keep.size == int(g["n_keep"]) and sha(keep.astype(np.int32)) == str(g["keep_sha"])
    # size-independent properties: kept set is an independent set, every dropped box has a kept suppressor
    kb = b[keep[:2000]]
    for i in (0, 17, 400):
        iou = pp.iou_one_vs_many(kb[i], kb[i + 1:])
        assert np.all(iou <= np.float32(0.45))


def test_k3_200k_80_classes(post):
    """BASELINE config 3: 200k total (box,class) candidates over 80 classes vs the C oracle"""
    b, o, c = cases.multiclass_case(200_000, 80, 2000, seed=3)
    R = nms_c.class_wise_nms(b, o, c, 0.45, 0.1)
    G = post.per_class_nms(b, o, c, 0.45, 0.1)
    assert all(np.array_equal(x, y) for x, y in zip(R, G))


@pytest.mark.parametrize("copies", [1, 40])
def test_iou_decisions_at_the_threshold_edge(post, copies):
    """Pairs whose IoU is EXACTLY the threshold, one ulp below and one ulp above it (integer boxes: IoU = 1/2, 1/3, 2/3, 3/4,
    1/5 as exact rationals rounded once), plus degenerate boxes (zero area, reversed corners, NaN / inf coordinates).  The
    warp route decides most pairs without the division (inter > thr*uni*(1 +- 2^-20)), the edge and the non-finite ones must
    fall back to the exact quotient: kept sets equal the C oracle on every route (copies = 40 puts the groups, shifted
    apart, into one segment of 440 boxes -> the CTA route)."""
    base = []
    for num, den in ((1, 2), (1, 3), (2, 3), (3, 4), (1, 5)):
        # box A = [0,0,den,1] (area den), box B = [0,0,num,1] inside it: IoU = num/den
        base.append(([0, 0, den, 1], [0, 0, num, 1], np.float32(num) / np.float32(den)))
    rng = np.random.default_rng(3)
    for (a, b, q) in base:
        for thr in (q, np.nextafter(q, np.float32(0)), np.nextafter(q, np.float32(1))):
            boxes, scores = [], []
            for c in range(copies):
                off = 50.0 * c
                boxes += [[a[0] + off, a[1], a[2] + off, a[3]], [b[0] + off, b[1], b[2] + off, b[3]]]
            boxes += [[3, 3, 3, 9], [9, 9, 5, 5], [np.nan, 0, 1, 1], [0, 0, np.inf, 1], [0, 0, 2, 2], [0, 0, 2, 2],
                      [1e30, 1e30, 3e38, 3e38], [-3e38, -3e38, 3e38, 3e38], [7, 7, 7, 7]]
            boxes = np.asarray(boxes, np.float32)
            scores = pp.make_tie_free_scores(len(boxes), rng)
            got = post.single_class_nms(boxes, scores, float(thr)).tolist()
            assert got == nms_c.greedy_nms(boxes, scores, float(thr)), (a, b, float(thr))
