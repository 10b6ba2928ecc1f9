"""Hardware probe (GPU): which UMMA descriptor convention reads row-shifted SWIZZLE_128B tiles correctly."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "object-detection-yolov3_b200"))
from yolo3_b200 import _lib  # noqa: E402
from yolo3_b200 import post_engine
from yolo3_b200._lib import check
eng = post_engine(0)
rng = np.random.default_rng(0)
a = rng.integers(-64, 64, (512, 64)).astype(np.float32)          # exactly representable in bf16
bits = (a.view(np.uint32) >> 16).astype(np.uint16)
shifts = np.array([0, 1, 2, 3, 7, 8, 9, 130], np.int32)
out = np.zeros((2, len(shifts), 128, 64), np.float32)
check(_lib.load_probe().y3_debug_umma_rowshift(eng.h, bits.ctypes.data, shifts.ctypes.data, len(shifts), out.ctypes.data), eng.h)
for v in range(2):
    for i, s in enumerate(shifts):
        ok = np.array_equal(out[v, i], a[s:s + 128])
        rows_ok = int((out[v, i] == a[s:s + 128]).all(axis=1).sum())
        print("base_offset=%s shift %3d: %s (%d/128 rows exact)" % ("0" if v == 0 else "(addr>>7)&7", s, "OK" if ok else "MISMATCH", rows_ok))
