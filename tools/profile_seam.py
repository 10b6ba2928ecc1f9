"""Cross-seam stage alone on the bench's own boxes (20000^2 image, one GPU) - measurement helper."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "object-detection-yolov3_b200"))
import bench  # noqa: E402
from yolo3_b200 import Engine  # noqa: E402

side = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
img = bench.synthetic_image(side, blobs=4000 * side * side // 400_000_000 + 10)
eng = Engine(bench.TILE + (1,), bench.NC, bench.ANCHORS, max_batch=256)
w = bench.bench_weights()
eng.load_weights(w)
bench.calibrate_heads(eng, w, eng.tiles_normalized(np.ascontiguousarray(img[:1024, :2048]), bench.TILE, 64, 0, 4))
dev = torch.device("cuda", 0)
img_dev = torch.from_numpy(img.view(np.int16)).to(dev).view(torch.uint16)
pred = eng.infer_tiled(img_dev, bench.TILE, bench.MIN_BOX, 64, bench.IOU_THR, bench.SCORE_THR, out_device=dev)
for i in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = eng.cross_seam_nms(pred, (side, side), bench.TILE, 64, 0.3)
    torch.cuda.synchronize()
    print("cross-seam stage: %d -> %d rows, %.3f ms" % (pred.shape[0], out.shape[0], 1e3 * (time.perf_counter() - t0)))

# the regime the sparse path is made for: detections of bounded size spread over the image (here 100 k boxes of 20-200 px)
rng = np.random.default_rng(5)
n = 100_000
cx, cy = rng.uniform(0, side, n), rng.uniform(0, side, n)
w, h = rng.uniform(20, 200, n), rng.uniform(20, 200, n)
rows = np.stack([np.clip(np.round(cx - w / 2), 0, side - 1), np.clip(np.round(cy - h / 2), 0, side - 1),
                 np.clip(np.round(cx + w / 2), 0, side - 1), np.clip(np.round(cy + h / 2), 0, side - 1),
                 rng.permutation(n).astype(np.float64) / n * 0.9 + 0.1, np.zeros(n)], 1)
rows_dev = torch.from_numpy(rows).to(dev)
for i in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = eng.cross_seam_nms(rows_dev, (side, side), bench.TILE, 64, 0.3)
    torch.cuda.synchronize()
    print("cross-seam stage, 100 k bounded boxes: %d -> %d rows, %.3f ms" % (n, out.shape[0], 1e3 * (time.perf_counter() - t0)))
