"""One steady K4 step (20000^2 uint16 image resident in HBM, 512^2 tiles, edge 64, batches of 256) after one warm-up step -
run under `ncu --metrics gpu__time_duration.sum` for the launch list of a bench step, or under `ncu --set full -k ...` for
one launch of a kernel (measurement helper; a number printed under ncu is never a bench value)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "object-detection-yolov3_b200"))
import bench  # noqa: E402
from yolo3_b200 import Engine  # noqa: E402

side = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
img = bench.synthetic_image(side, blobs=4000 * side * side // 400_000_000 + 10)
eng = Engine(bench.TILE + (1,), bench.NC, bench.ANCHORS, max_batch=batch)
w = bench.bench_weights()
eng.load_weights(w)
bench.calibrate_heads(eng, w, eng.tiles_normalized(np.ascontiguousarray(img[:1024, :2048]), bench.TILE, 64, 0, 4))
dev = torch.device("cuda", 0)
img_dev = torch.from_numpy(img.view(np.int16)).to(dev).view(torch.uint16)
eng.infer_tiled(img_dev, bench.TILE, bench.MIN_BOX, 64, bench.IOU_THR, bench.SCORE_THR, out_device=dev)
k0 = eng.timings()["kernels_launched"]
print("MEASURED STEP")
pred = eng.infer_tiled(img_dev, bench.TILE, bench.MIN_BOX, 64, bench.IOU_THR, bench.SCORE_THR, out_device=dev)
t = eng.timings()
print(tuple(pred.shape), "launches in the step:", t["kernels_launched"] - k0, t)
