"""NumPy restatement of the reference's tile slicing / normalisation / stitching
(ORACLE - tests only).

Follows (does not copy) /root/reference:
  zscore                 imagereader.py:34-46
  tile_plan / cut_tiles  inference_tiled.py:29-100   (convert_image_to_tiles)
  tiled_inference        inference_tiled.py:185-310  (inference_image_tiled)

Bug-compatible behaviours that are kept on purpose:
  Q12  the recorded origin of a border tile is clamped to 0 although the tile
       was reflect-padded on that side (tile-local -> global shifted by +r);
  Q13  seams are resolved by centre ownership, there is no cross-tile NMS;
  Q14  every tile is normalised with its own mean / population std, and
       std <= 1.0 means "subtract the mean only".
"""
import numpy as np

from . import postproc_np as pp

F32 = np.float32


def zscore(a):
    """imagereader.py:34-46 (fp32 statistics, population std)."""
    a = np.asarray(a).astype(F32)
    sd = np.std(a)
    mu = np.mean(a)
    return (a - mu) if sd <= 1.0 else (a - mu) / sd


def tile_plan(height, width, tile_size, edge_range):
    """Geometry of inference_tiled.py:29-100 without touching pixels.

    Returns a list of dicts in the reference's tile order (rows, then columns):
      y0,y1,x0,x1   clamped crop [y0:y1, x0:x1] in the image
      pad           ((top, bottom), (left, right)) reflect padding
      rec_x, rec_y  the origin the reference RECORDS (clamped -> Q12)
    and the per-axis radius actually used.
    """
    th, tw = int(tile_size[0]), int(tile_size[1])
    assert th % 32 == 0 and tw % 32 == 0
    ry = 0 if th >= height else int(edge_range)
    rx = 0 if tw >= width else int(edge_range)
    assert ry % 32 == 0 and rx % 32 == 0
    zy, zx = th - 2 * ry, tw - 2 * rx
    plan = []
    for i in range(0, height, zy):
        for j in range(0, width, zx):
            ys, ye = i - ry, i + zy + ry
            xs, xe = j - rx, j + zx + rx
            top, left = max(0, -ys), max(0, -xs)
            bot, right = max(0, ye - height), max(0, xe - width)
            ys, xs = max(ys, 0), max(xs, 0)
            ye, xe = min(ye, height), min(xe, width)
            plan.append(dict(y0=ys, y1=ye, x0=xs, x1=xe, pad=((top, bot), (left, right)),
                             rec_x=xs, rec_y=ys))
    return plan, (ry, rx)


def cut_tiles(img, tile_size, edge_range):
    """inference_tiled.py:29-100.  img is HxWxC.  -> (tiles, xs, ys)."""
    plan, _ = tile_plan(img.shape[0], img.shape[1], tile_size, edge_range)
    tiles, xs, ys = [], [], []
    for p in plan:
        t = img[p["y0"]:p["y1"], p["x0"]:p["x1"]]
        (pt, pb), (pl, pr) = p["pad"]
        if pt or pb or pl or pr:
            t = np.pad(t, ((pt, pb), (pl, pr), (0, 0)), mode="reflect")
        tiles.append(t)
        xs.append(p["rec_x"])
        ys.append(p["rec_y"])
    return tiles, xs, ys


def ghost_band_mask(boxes, org_x, org_y, img_hw, tile_size, edge_range):
    """inference_tiled.py:235-254: True where the box centre is NOT owned by this tile."""
    r = edge_range
    cx = (boxes[:, 2] + boxes[:, 0]) / 2.0          # fp32
    cy = (boxes[:, 3] + boxes[:, 1]) / 2.0
    gx = cx + F32(org_x)
    gy = cy + F32(org_y)
    bad = (gy > r) & (cy < r)
    bad |= (gy <= img_hw[0] - r) & (cy >= tile_size[0] - r)
    bad |= (gx > r) & (cx < r)
    bad |= (gx <= img_hw[1] - r) & (cx >= tile_size[1] - r)
    return bad


def finish_boxes(boxes, scores, labels, img_hw):
    """inference_tiled.py:278-301 on concatenated fp32 global boxes."""
    H, W = int(img_hw[0]), int(img_hw[1])
    b = np.round(boxes).astype(np.int32)
    cx = (b[:, 2] + b[:, 0]) / 2.0
    cy = (b[:, 3] + b[:, 1]) / 2.0
    ok = ~((cx < 0) | (cx >= W) | (cy < 0) | (cy >= H))
    b, scores, labels = b[ok], scores[ok], labels[ok]
    b[:, 0] = np.clip(b[:, 0], 0, W - 1)
    b[:, 2] = np.clip(b[:, 2], 0, W - 1)
    b[:, 1] = np.clip(b[:, 1], 0, H - 1)
    b[:, 3] = np.clip(b[:, 3], 0, H - 1)
    return b, scores, labels


def tiled_inference(model_fn, img, tile_size, min_roi_size, edge_range=96,
                    iou_threshold=0.3, score_threshold=0.1, nms_fn=pp.greedy_nms):
    """inference_tiled.py:185-310.

    model_fn(batch[1,C,H,W] f32) -> ndarray[1,N,5+NC] f32.
    Returns float64 [n,6] rows (x0,y0,x1,y1,score,label) in the reference's order.
    """
    H, W = img.shape[0], img.shape[1]
    tiles, xs, ys = cut_tiles(img, tile_size, edge_range)
    _, (ry, rx) = tile_plan(H, W, tile_size, edge_range)
    acc_b, acc_s, acc_l = [], [], []
    for t, ox, oy in zip(tiles, xs, ys):
        x = zscore(t.astype(F32)).transpose(2, 0, 1)[None]
        det = np.asarray(model_fn(np.ascontiguousarray(x)))[0]
        det = pp.drop_small(det, min_roi_size)
        b, s, l = pp.class_wise_nms(det[:, 0:4], det[:, 4:5], det[:, 5:], iou_threshold,
                                    score_threshold, nms_fn=nms_fn)
        if b is None:
            continue
        # the reference tests both axes against the single constant EDGE_EFFECT_RANGE
        bad = ghost_band_mask(b, ox, oy, (H, W), tile_size, edge_range)
        b, s, l = b[~bad].copy(), s[~bad], l[~bad]
        if b.shape[0] == 0:
            continue
        b[:, 0] += ox
        b[:, 2] += ox
        b[:, 1] += oy
        b[:, 3] += oy
        acc_b.append(b)
        acc_s.append(s)
        acc_l.append(l)
    if not acc_b:
        return np.zeros((0, 6), dtype=np.float64)
    b, s, l = finish_boxes(np.concatenate(acc_b), np.concatenate(acc_s), np.concatenate(acc_l), (H, W))
    return np.concatenate((b.astype(np.float64), s.astype(np.float64)[:, None],
                           l.astype(np.float64)[:, None]), axis=1)


def tiled_inference_single_tile(model_fn, img, tile_size, min_roi_size, edge_range, tile_index,
                                iou_threshold=0.3, score_threshold=0.1, nms_fn=pp.greedy_nms):
    """The rows of tiled_inference() that come from ONE tile (tile-major output => the full result is
    the concatenation over tiles).  Used to check the rank-sharded path on the CPU."""
    H, W = img.shape[0], img.shape[1]
    tiles, xs, ys = cut_tiles(img, tile_size, edge_range)
    t, ox, oy = tiles[tile_index], xs[tile_index], ys[tile_index]
    x = zscore(t.astype(F32)).transpose(2, 0, 1)[None]
    det = np.asarray(model_fn(np.ascontiguousarray(x)))[0]
    det = pp.drop_small(det, min_roi_size)
    b, s, l = pp.class_wise_nms(det[:, 0:4], det[:, 4:5], det[:, 5:], iou_threshold, score_threshold, nms_fn=nms_fn)
    if b is None:
        return np.zeros((0, 6), np.float64)
    bad = ghost_band_mask(b, ox, oy, (H, W), tile_size, edge_range)
    b, s, l = b[~bad].copy(), s[~bad], l[~bad]
    if b.shape[0] == 0:
        return np.zeros((0, 6), np.float64)
    b[:, 0] += ox
    b[:, 2] += ox
    b[:, 1] += oy
    b[:, 3] += oy
    b, s, l = finish_boxes(b, s, l, (H, W))
    return np.concatenate((b.astype(np.float64), s.astype(np.float64)[:, None], l.astype(np.float64)[:, None]), axis=1)


def seam_candidates(pred, img_hw, tile_size, edge_range):
    """Rows of the final [n,6] result whose inclusive integer extent straddles a zone boundary."""
    pred = np.asarray(pred)
    cand = np.zeros(pred.shape[0], bool)
    for axis, (lo, hi) in enumerate(((1, 3), (0, 2))):
        if int(tile_size[axis]) >= int(img_hw[axis]):
            continue
        zone = int(tile_size[axis]) - 2 * int(edge_range)
        cand |= (pred[:, lo] // zone) != (pred[:, hi] // zone)
    return cand


def cross_seam_nms(pred, img_hw, tile_size, edge_range=96, iou_threshold=0.3):
    """Optional stage that is NOT in the reference (north_star's cross-seam NMS): greedy per-class NMS
    (the pinned greedy_nms, tie rule score desc / row asc) among the seam candidates only."""
    pred = np.asarray(pred, np.float64)
    cand = seam_candidates(pred, img_hw, tile_size, edge_range)
    drop = np.zeros(pred.shape[0], bool)
    rows = np.nonzero(cand)[0]
    for c in np.unique(pred[rows, 5]):
        r = rows[pred[rows, 5] == c]
        kept = pp.greedy_nms(pred[r, 0:4].astype(F32), pred[r, 4].astype(F32), iou_threshold)
        gone = np.ones(r.size, bool)
        gone[kept] = False
        drop[r[gone]] = True
    return pred[~drop]
