"""Drop-in for the reference's bbox_utils.py hot functions - same names, arguments and return
types, computed by libyolo3_b200.so on the GPU (no NumPy fallback).

  compute_iou / single_class_nms / per_class_nms / filter_small_boxes
      reference bbox_utils.py:200-281 -> y3_compute_iou / y3_single_class_nms / y3_per_class_nms /
      y3_filter_small (include/yolo3_b200.h)
  write_boxes_from_xywhc / write_boxes_from_ltrbpc
      reference bbox_utils.py:47-62, 284-300 - CSV glue, plain Python.
Only the functions the two inference scripts use are provided (the draw /
merge / CSV-load helpers are outside the hot path).
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
