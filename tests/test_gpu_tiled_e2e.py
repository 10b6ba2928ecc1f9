"""y3_infer_tiled (a11-a17 with the real network) through the C ABI.

Exactness split (DESIGN.md "parity"):
  * everything around the network is integer / fp32-exact: with the GPU's OWN decoded boxes injected as
    the model, the reference-pinned oracle pipeline must reproduce y3_infer_tiled's [n,6] rows bit for bit;
  * the network itself is tolerance-matched (heads 2e-2) in test_gpu_net.py.
"""
import numpy as np
import pytest

from oracle import cases, nms_c, postproc_np as pp, tiling_np as tl

pytestmark = pytest.mark.gpu

TILE = (256, 256)
NC = 2


def standardised_engine(img, edge, max_batch, obj_bias=-4.0, seed=0):
    """random-init network whose detection channels are standardised on the GPU path itself (as bench.py)"""
    from yolo3_b200 import Engine, weights
    anchors = [(32, 32), (64, 64), (128, 128)]
    w = weights.random_init(1, NC, 3, seed=seed, randomize_bn=True)
    eng = Engine(TILE + (1,), NC, anchors, max_batch=max_batch)
    eng.load_weights(w)
    from yolo3_b200 import tile_count
    sample = eng.tiles_normalized(img, TILE, edge, 0, min(2, tile_count(img.shape[0], img.shape[1], TILE, edge)))
    E = 5 + NC
    upd = {}
    for i, h in enumerate(eng.forward_heads(sample)):
        k = "feature_map_%d" % (i + 1)
        mu, sd = h.mean(axis=(0, 2, 3)), np.maximum(h.std(axis=(0, 2, 3)), 1e-12)
        want = np.zeros(3 * E)
        want.reshape(3, E)[:, 4] = obj_bias
        upd[k + "/kernel"] = (w[k + "/kernel"] / sd[None, None, None, :]).astype(np.float32)
        upd[k + "/bias"] = (want - mu / sd).astype(np.float32)
    eng.load_weights(upd)
    return eng


@pytest.mark.parametrize("shape,edge,batch", [((700, 900, 1), 64, 4), ((520, 1100, 1), 32, 3), ((200, 230, 1), 96, 2)])
def test_infer_tiled_equals_oracle_pipeline_on_gpu_boxes(shape, edge, batch):
    img = cases.synthetic_image(*shape, np.uint16, seed=shape[0], blobs=20)
    eng = standardised_engine(img, edge, batch)
    pred = eng.infer_tiled(img, TILE, 24, edge_range=edge)
    tiles = eng.tiles_normalized(img, TILE, edge)
    dets = np.concatenate([eng.forward_boxes(tiles[i:i + batch]) for i in range(0, len(tiles), batch)])
    it = iter(range(len(tiles)))
    want = tl.tiled_inference(lambda x: dets[next(it)][None], img, TILE, 24, edge_range=edge, nms_fn=nms_c.greedy_nms)
    print("tiles", len(tiles), "boxes", pred.shape, want.shape)
    assert want.shape[0] > (10 if len(tiles) > 1 else 0)
    assert pred.dtype == np.float64 and np.array_equal(pred, want)
    # sharded exactly like the multi-GPU path: 3 tile ranges, concatenated in rank order
    from yolo3_b200 import shard_range
    parts = [eng.infer_tiled(img, TILE, 24, edge_range=edge, tile_first=f, tile_count=c)
             for f, c in (shard_range(len(tiles), r, 3) for r in range(3))]
    assert np.array_equal(np.concatenate(parts), want)


def test_cross_seam_nms_matches_oracle():
    """optional stage (not in the reference): GPU NMS among the seam-straddling final boxes == oracle stage"""
    rng = np.random.default_rng(11)
    H, W, tile, edge = 1500, 1700, (512, 512), 96
    n = 6000
    cx, cy = rng.uniform(0, W, n), rng.uniform(0, H, n)
    w, h = rng.uniform(20, 120, n), rng.uniform(20, 120, n)
    pred = np.stack([np.clip(np.round(cx - w / 2), 0, W - 1), np.clip(np.round(cy - h / 2), 0, H - 1),
                     np.clip(np.round(cx + w / 2), 0, W - 1), np.clip(np.round(cy + h / 2), 0, H - 1),
                     pp.make_tie_free_scores(n, rng).astype(np.float64), rng.integers(0, 3, n).astype(np.float64)], axis=1)
    from yolo3_b200 import post_engine, seam_candidates
    cand = seam_candidates(pred, (H, W), tile, edge)
    assert np.array_equal(cand, tl.seam_candidates(pred, (H, W), tile, edge)) and 100 < cand.sum() < n
    got = post_engine().cross_seam_nms(pred, (H, W), tile, edge, 0.3)
    want = tl.cross_seam_nms(pred, (H, W), tile, edge, 0.3)
    assert got.shape[0] < n and np.array_equal(got, want)
    # rows that do not straddle a seam are never touched
    assert np.array_equal(got[~seam_candidates(got, (H, W), tile, edge)], pred[~cand])
