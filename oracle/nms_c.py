"""ctypes wrapper of oracle/nms_oracle.c (ORACLE - tests only)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "libnms_oracle.so"])


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(_HERE, "libnms_oracle.so")
        if not os.path.exists(so):
            build()
        L = ctypes.CDLL(so)
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int32)
        L.y3o_single_class_nms.restype = ctypes.c_int64
        L.y3o_single_class_nms.argtypes = [fp, fp, ctypes.c_int64, ctypes.c_float, ip]
        L.y3o_per_class_nms.restype = ctypes.c_int64
        L.y3o_per_class_nms.argtypes = [fp, fp, fp, ctypes.c_int64, ctypes.c_int32, ctypes.c_float,
                                        ctypes.c_float, fp, fp, ip, ctypes.c_int64]
        _lib = L
    return _lib


def _f(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _i(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def greedy_nms(boxes, scores, iou_threshold):
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    m = boxes.shape[0]
    keep = np.empty(max(m, 1), dtype=np.int32)
    k = lib().y3o_single_class_nms(_f(boxes), _f(scores), m, np.float32(iou_threshold), _i(keep))
    return keep[:k].tolist()


def class_wise_nms(boxes, objectness, class_probs, iou_threshold=0.3, score_threshold=0.1):
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    obj = np.ascontiguousarray(objectness, dtype=np.float32).reshape(-1)
    cls = np.ascontiguousarray(class_probs, dtype=np.float32)
    n, nc = cls.shape
    cap = max(n * nc, 1)
    ob = np.empty((cap, 4), np.float32)
    os_ = np.empty(cap, np.float32)
    ol = np.empty(cap, np.int32)
    k = lib().y3o_per_class_nms(_f(boxes), _f(obj), _f(cls), n, nc, np.float32(iou_threshold),
                                np.float32(score_threshold), _f(ob), _f(os_), _i(ol), cap)
    assert k >= 0
    if k == 0:
        return None, None, None
    return ob[:k].copy(), os_[:k].copy(), ol[:k].copy()
