"""Drop-in for the reference's bbox_utils.py hot functions - same names, arguments and return
types, computed by libyolo3_b200.so on the GPU (no NumPy fallback).

  compute_iou / single_class_nms / per_class_nms / filter_small_boxes
      reference bbox_utils.py:200-281 -> y3_compute_iou / y3_single_class_nms / y3_per_class_nms /
      y3_filter_small (include/yolo3_b200.h)
  write_boxes_from_xywhc / write_boxes_from_ltrbpc
      reference bbox_utils.py:47-62, 284-300 - CSV glue, plain Python.
The merge / CSV-load helpers (outside the hot path) are at the end of the file; the matplotlib draw helper is not
provided.
"""
import numpy as np

from yolo3_b200 import post_engine


def compute_iou(box, boxes, box_area=None, boxes_area=None):
    # the areas are recomputed on the device exactly as the reference does when they are None;
    # passing them is accepted for signature compatibility (they are the same fp32 products).
    return post_engine().compute_iou(np.asarray(box, np.float32), np.asarray(boxes, np.float32))


def single_class_nms(boxes, scores, iou_threshold):
    return [int(i) for i in post_engine().single_class_nms(boxes, scores, iou_threshold)]


def per_class_nms(boxes, objectness, class_probs, iou_threshold=0.3, score_threshold=0.1):
    return post_engine().per_class_nms(boxes, objectness, class_probs, iou_threshold, score_threshold)


def filter_small_boxes(boxes, min_roi_size):
    return post_engine().filter_small(boxes, min_roi_size)


def write_boxes_from_xywhc(boxes, csv_filename):
    with open(csv_filename, "w") as fh:
        fh.write("X,Y,W,H,C\n")
        for row in np.asarray(boxes):
            fh.write("%d,%d,%d,%d,%d\n" % tuple(int(v) for v in row[:5]))


def write_boxes_from_ltrbpc(boxes, csv_filename):
    with open(csv_filename, "w") as fh:
        fh.write("X,Y,W,H,P,C\n")
        for row in np.asarray(boxes):
            x, y = int(row[0]), int(row[1])
            fh.write("{:d},{:d},{:d},{:d},{:f},{:d}\n".format(x, y, int(row[2] - x + 1), int(row[3] - y + 1),
                                                            row[4], int(row[5])))


# ------------------------------------------------------------------------------------------------
# Helpers of the reference module that are off the inference hot path (bbox_utils.py:65-197): CSV glue in
# plain Python, and the greedy box-merge utility with its IoU evaluations on the GPU.
def write_boxes_from_ltrbc(boxes, csv_filename):
    """rows x0, y0, x1, y1, class -> X,Y,W,H,C with W = x1 - x0 + 1 (bbox_utils.py:65-80)."""
    with open(csv_filename, "w") as fh:
        fh.write("X,Y,W,H,C\n")
        for row in np.asarray(boxes):
            x, y = int(row[0]), int(row[1])
            fh.write("%d,%d,%d,%d,%d\n" % (x, y, int(row[2]) - x + 1, int(row[3]) - y + 1, int(row[4])))


def _read_xywhc(filepath):
    import csv
    import os
    rows = []
    if os.path.exists(filepath):
        with open(filepath) as fh:
            for rec in csv.DictReader(fh, skipinitialspace=True):
                rows.append([int(rec[k]) for k in ("X", "Y", "W", "H", "C")])
    return np.asarray(rows, dtype=np.float64).reshape(-1, 5)


def load_boxes_to_xywhc(filepath):
    """CSV written by write_boxes_from_xywhc -> float array [n, 5] (bbox_utils.py:107-125); missing file -> [0, 5]."""
    return _read_xywhc(filepath)


def load_boxes_to_ltrbc(filepath):
    """as above with W, H converted to inclusive end coordinates (bbox_utils.py:83-104)."""
    a = _read_xywhc(filepath)
    a[:, 2] = a[:, 0] + a[:, 2] - 1
    a[:, 3] = a[:, 1] + a[:, 3] - 1
    return a


def box_union(boxes, weights):
    """bounding box of `boxes` ([1, 4]) and the mean of `weights` (bbox_utils.py:128-135)."""
    boxes = np.asarray(boxes)
    bb = np.array([[boxes[:, 0].min(), boxes[:, 1].min(), boxes[:, 2].max(), boxes[:, 3].max()]], dtype=np.float64)
    return bb, np.mean(weights)


def union_all_overlapping_bb(boxes, scores, minimum_iou_for_merge=0):
    """Greedy merge of overlapping boxes (bbox_utils.py:138-197): walk the boxes in descending score order; the
    head box absorbs every remaining box whose IoU with it exceeds the threshold (union box, mean score) and goes
    to the back of the queue; stop after a full pass without a merge.  Mutates `boxes` / `scores` in place like
    the reference and returns the surviving rows in queue order.  The IoU rows are evaluated by y3_compute_iou."""
    if len(scores) <= 1:
        return boxes, scores
    if boxes.dtype.kind == "i":
        boxes = boxes.astype("float")
    queue = scores.argsort()[::-1].tolist()
    quiet = 0
    eng = post_engine()
    while len(queue) > 1 and quiet <= len(queue):
        head = queue.pop(0)
        ious = eng.compute_iou(np.asarray(boxes[head], np.float32), np.asarray(boxes[queue], np.float32))
        hit = np.nonzero(ious > minimum_iou_for_merge)[0]
        if hit.size:
            quiet = 0
            members = np.append(np.asarray(queue)[hit], head)
            bb, w = box_union(boxes[members], scores[members])
            boxes[head, 0:4] = bb[0]
            scores[head] = w
            gone = set(hit.tolist())
            queue = [v for i, v in enumerate(queue) if i not in gone]
        else:
            quiet += 1
        queue.append(head)
    keep = np.array(queue)
    return boxes[keep, :], scores[keep]
