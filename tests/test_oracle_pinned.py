"""Container-only: the restatements against the reference's code executed verbatim on
fresh random inputs (beyond the committed fixtures).  Skips where /root/reference is absent."""
import contextlib
import io

import numpy as np
import pytest

from oracle import cases, nms_c, postproc_np as pp, tiling_np as tl


@pytest.mark.parametrize("seed", [101, 102, 103])
def test_nms_vs_reference(reference, seed):
    bu = reference["bbox_utils"]
    b, s = cases.nms_case(2500, 600 + 100 * (seed % 3), seed, wh=(10, 200))
    for thr in (0.3, 0.45):
        ref = [int(i) for i in bu.single_class_nms(b, s, thr)]
        assert pp.greedy_nms(b, s, thr) == ref
        assert nms_c.greedy_nms(b, s, thr) == ref


def test_per_class_vs_reference(reference):
    bu = reference["bbox_utils"]
    b, o, c = cases.multiclass_case(3000, 7, 500, seed=77, dominant_only=False)
    R = bu.per_class_nms(b, o, c, 0.3, 0.1)
    for got in (pp.class_wise_nms(b, o, c, 0.3, 0.1), nms_c.class_wise_nms(b, o, c, 0.3, 0.1)):
        assert all(np.array_equal(x, y) for x, y in zip(R, got))


@pytest.mark.parametrize("edge,tile,shape", [(96, (512, 512), (1111, 777, 1)), (32, (128, 160), (500, 333, 3)),
                                             (64, (512, 512), (512, 2000, 1))])
def test_tiled_vs_reference(reference, edge, tile, shape):
    it = reference["inference_tiled"]
    img = cases.synthetic_image(*shape, np.uint16 if shape[2] == 1 else np.uint8, seed=edge + shape[0])
    fake = cases.FakeDetector(500, 2, tile, seed=edge)
    old = it.EDGE_EFFECT_RANGE
    try:
        it.EDGE_EFFECT_RANGE = edge
        with contextlib.redirect_stdout(io.StringIO()):
            ref = it.inference_image_tiled(fake, img, list(tile), 20)
        t_ref, xs, ys = it.convert_image_to_tiles(img, list(tile))
    finally:
        it.EDGE_EFFECT_RANGE = old
    got = tl.tiled_inference(fake, img, tile, 20, edge_range=edge)
    assert np.array_equal(ref, got) and ref.shape[0] > 0
    t_got, xg, yg = tl.cut_tiles(img, tile, edge)
    assert xs == xg and ys == yg and all(np.array_equal(a, b) for a, b in zip(t_ref, t_got))


def test_cross_seam_stage_only_touches_seam_boxes():
    """oracle of the optional cross-seam NMS stage (not in the reference): a duplicate pair straddling a seam
    loses its lower-scored box, identical duplicates away from any seam are left alone"""
    from oracle import tiling_np as tl
    H, W, tile, edge = 1000, 1000, (512, 512), 96           # zone 320 -> seams at 320, 640, 960
    pred = np.array([[300, 100, 340, 140, 0.9, 0], [301, 101, 341, 141, 0.8, 0],       # straddle x = 320: duplicate
                     [300, 100, 340, 140, 0.7, 1],                                       # other class: kept
                     [100, 100, 140, 140, 0.9, 0], [101, 101, 141, 141, 0.8, 0],         # no seam: untouched
                     [100, 630, 140, 650, 0.6, 0], [100, 631, 141, 651, 0.5, 0]], np.float64)   # straddle y = 640
    cand = tl.seam_candidates(pred, (H, W), tile, edge)
    assert cand.tolist() == [True, True, True, False, False, True, True]
    out = tl.cross_seam_nms(pred, (H, W), tile, edge, 0.3)
    assert np.array_equal(out, pred[[0, 2, 3, 4, 5]])
    # an axis that is not tiled has no seams
    assert not tl.seam_candidates(pred, (400, 1000), tile, edge)[5:].any()
