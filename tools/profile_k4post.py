"""K4-like tiled run on a small image (6000^2 -> 256 tiles) - run under `ncu --metrics gpu__time_duration.sum` to list
the per-batch post-processing launches without contention from a following batch (measurement helper)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "object-detection-yolov3_b200"))
import bench  # noqa: E402
from yolo3_b200 import Engine  # noqa: E402

side = 6000
img = bench.synthetic_image(side, blobs=400)
eng = Engine(bench.TILE + (1,), bench.NC, bench.ANCHORS, max_batch=256)
w = bench.bench_weights()
eng.load_weights(w)
bench.calibrate_heads(eng, w, eng.tiles_normalized(np.ascontiguousarray(img[:1024, :2048]), bench.TILE, 64, 0, 4))
eng.infer_tiled(img, bench.TILE, bench.MIN_BOX, 64, bench.IOU_THR, bench.SCORE_THR)
print("MEASURED CALL")
pred = eng.infer_tiled(img, bench.TILE, bench.MIN_BOX, 64, bench.IOU_THR, bench.SCORE_THR)
print(pred.shape, eng.timings())
