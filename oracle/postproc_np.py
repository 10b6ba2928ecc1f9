"""NumPy restatement of the reference's box post-processing (ORACLE - tests only).

Follows (does not copy) /root/reference:
  compute_iou           bbox_utils.py:200-214  (= inference_tiled.py:103-117)
  single_class_nms      bbox_utils.py:217-237  (= inference_tiled.py:120-140)
  per_class_nms         bbox_utils.py:240-271  (= inference_tiled.py:143-173)
  filter_small_boxes    bbox_utils.py:274-281  (= inference_tiled.py:176-182)

All arithmetic is fp32, in the reference's operand order, because the GPU
path must reproduce the kept-index set bit for bit:
  inter = max(yb - yt, 0) * max(xr - xl, 0)
  union = (area_box + area_boxes) - inter
  iou   = inter / union            (0/0 -> NaN -> "iou <= thr" False -> suppressed)
  score = sqrt(cls * obj) ; candidate iff score >= float32(score_thr)
  survivor iff iou <= float32(iou_thr)

Tie rule: the reference sorts with `scores.argsort()[::-1]` (unstable, SURVEY
Q11).  This restatement sorts by (score desc, index asc) - the GPU rule - and
is only compared with the reference on tie-free scores.
"""
import numpy as np

F32 = np.float32


def box_areas(boxes):
    boxes = np.asarray(boxes, dtype=F32)
    return (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])


def iou_one_vs_many(box, boxes, box_area=None, boxes_area=None):
    """bbox_utils.py:200-214."""
    boxes = np.asarray(boxes, dtype=F32)
    box = np.asarray(box, dtype=F32)
    xl = np.maximum(box[0], boxes[:, 0])
    yt = np.maximum(box[1], boxes[:, 1])
    xr = np.minimum(box[2], boxes[:, 2])
    yb = np.minimum(box[3], boxes[:, 3])
    inter = np.maximum(yb - yt, F32(0)) * np.maximum(xr - xl, F32(0))
    if box_area is None:
        box_area = (box[2] - box[0]) * (box[3] - box[1])
    if boxes_area is None:
        boxes_area = box_areas(boxes)
    with np.errstate(divide="ignore", invalid="ignore"):
        return inter / ((F32(box_area) + boxes_area) - inter)


def order_desc(scores):
    """score descending, index ascending on ties (documented GPU tie rule)."""
    scores = np.asarray(scores, dtype=F32)
    idx = np.arange(scores.shape[0])
    return np.lexsort((idx, -scores.astype(np.float64)))


def greedy_nms(boxes, scores, iou_threshold):
    """bbox_utils.py:217-237.  Returns kept indices (list[int]) in pick order."""
    boxes = np.asarray(boxes, dtype=F32)
    thr = F32(iou_threshold)
    areas = box_areas(boxes)
    alive = order_desc(scores)
    kept = []
    while alive.size:
        head, alive = alive[0], alive[1:]
        kept.append(int(head))
        if not alive.size:
            break
        iou = iou_one_vs_many(boxes[head], boxes[alive], areas[head], areas[alive])
        alive = alive[iou <= thr]
    return kept


def blended_scores(objectness, class_probs):
    """bbox_utils.py:244-245: sqrt(class_prob * objectness), fp32."""
    obj = np.asarray(objectness, dtype=F32).reshape(-1, 1)
    cls = np.asarray(class_probs, dtype=F32)
    return np.sqrt(cls * obj)


def class_wise_nms(boxes, objectness, class_probs, iou_threshold=0.3, score_threshold=0.1,
                   nms_fn=greedy_nms):
    """bbox_utils.py:240-271.  (boxes[k,4] f32, scores[k] f32, labels[k] i32) or (None,)*3."""
    boxes = np.asarray(boxes, dtype=F32)
    sc = blended_scores(objectness, class_probs)
    out_b, out_s, out_l = [], [], []
    for c in range(sc.shape[1]):
        sel = np.nonzero(sc[:, c] >= F32(score_threshold))[0]
        if sel.size == 0:
            continue
        cb, cs = boxes[sel], sc[sel, c]
        keep = np.asarray(nms_fn(cb, cs, iou_threshold), dtype=np.int64)
        out_b.append(cb[keep])
        out_s.append(cs[keep])
        out_l.append(np.full(keep.shape[0], c, dtype=np.int32))
    if not out_b:
        return None, None, None
    return np.concatenate(out_b, 0), np.concatenate(out_s, 0), np.concatenate(out_l, 0)


def drop_small(boxes, min_size):
    """bbox_utils.py:274-281: keep rows with width > min AND height > min (strict, fp32)."""
    boxes = np.asarray(boxes)
    w = boxes[:, 2] - boxes[:, 0]
    h = boxes[:, 3] - boxes[:, 1]
    return boxes[(w > min_size) & (h > min_size)]


def make_tie_free_scores(n, rng, lo=0.05, hi=0.999):
    """n distinct fp32 values in (lo, hi), shuffled - so SURVEY Q11 cannot bite."""
    ladder = np.unique(np.linspace(lo, hi, 4 * n + 17).astype(F32))
    assert ladder.size >= n
    pick = rng.choice(ladder.size, size=n, replace=False)
    return ladder[pick].astype(F32)
