"""Measurement helper (GPU): per-head max|a-b|/max|b| vs the fp32 oracle over weight seeds.
    python tests/head_error_sweep.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "object-detection-yolov3_b200")):
    sys.path.insert(0, p)
from oracle import model_torch as mt  # noqa: E402
from yolo3_b200 import Engine  # noqa: E402

CFGS = [((512, 512, 1), 1, [(64, 384), (384, 64)]), ((512, 512, 1), 1, None), ((416, 416, 3), 80, None)]
for img_size, nc, anchors in CFGS:
    for seed in (0, 1, 2, 3, 4, 5):
        W = mt.init_weights(img_size[2], nc, len(anchors or mt.DEFAULT_ANCHORS), seed=seed, randomize_bn=True)
        eng = Engine(img_size, nc, anchors, max_batch=1)
        eng.load_weights({k: v.numpy() for k, v in W.items()})
        x = torch.randn(1, img_size[2], img_size[0], img_size[1], generator=torch.Generator().manual_seed(2))
        want = mt.OracleNet(W, img_size, nc, anchors).feature_maps(x)
        got = eng.forward_heads(x.numpy())
        print(img_size, nc, "A=%d" % len(anchors or mt.DEFAULT_ANCHORS), "seed", seed,
              ["%.4f" % mt.heads_rel_err(a, b.numpy()) for a, b in zip(got, want)], flush=True)
        eng.close()
