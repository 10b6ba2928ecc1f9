"""Helper of test_gpu_variants.py (runs in a subprocess because the A/B switches are read once per process):
prints the per-head relative errors of two small configurations against the oracle as JSON."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "object-detection-yolov3_b200"))
from oracle import model_torch as mt  # noqa: E402
from yolo3_b200 import Engine  # noqa: E402

out = {}
for img_size, nc, B in (((160, 224, 3), 3, 3), ((256, 256, 1), 1, 3)):
    W = mt.init_weights(img_size[2], nc, 3, seed=1, randomize_bn=True)
    eng = Engine(img_size, nc, None, max_batch=B)
    eng.load_weights({k: v.numpy() for k, v in W.items()})
    ora = mt.OracleNet(W, img_size, nc, None)
    x = torch.randn(B, img_size[2], img_size[0], img_size[1], generator=torch.Generator().manual_seed(3))
    want = ora.feature_maps(x)
    got = eng.forward_heads(x.numpy())
    out["%dx%dx%d" % img_size] = [float(mt.heads_rel_err(a, b.numpy())) for a, b in zip(got, want)]
print("RESULT " + json.dumps(out))
