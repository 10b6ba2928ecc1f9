"""Network forward (a1-a10) through the C ABI vs the torch-CPU restatement (parity unpinned: no TF).
Tolerance from north_star: per-head max|a-b| / max|b| <= 2e-2 against the fp32 oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import model_torch as mt, nms_c, postproc_np as pp

pytestmark = pytest.mark.gpu
HEAD_TOL = 2e-2


def make(img_size, nc, anchors=None, max_batch=2, seed=0, **kw):
    from yolo3_b200 import Engine
    W = mt.init_weights(img_size[2], nc, len(anchors or mt.DEFAULT_ANCHORS), seed=seed, randomize_bn=True, **kw)
    eng = Engine(img_size, nc, anchors, max_batch=max_batch)
    eng.load_weights({k: v.numpy() for k, v in W.items()})
    return eng, mt.OracleNet(W, img_size, nc, anchors)


def test_layers_small_net():
    """every layer of a small network against the oracle's trace - localises any kernel fault"""
    os.environ["Y3_DEBUG_NO_REUSE"] = "1"
    try:
        eng, ora = make((96, 128, 3), 2, max_batch=2)
    finally:
        del os.environ["Y3_DEBUG_NO_REUSE"]
    x = torch.randn(2, 3, 96, 128, generator=torch.Generator().manual_seed(1))
    ora.trace = {}
    want = ora.feature_maps(x)
    got = eng.forward_heads(x.numpy())
    report = []
    for name, ref in ora.trace.items():
        out = eng.debug_layer_output(name, 2)
        err = mt.heads_rel_err(out, ref.numpy())
        report.append((name, tuple(ref.shape), err))
    bad = [r for r in report if not r[2] < 3e-2]
    print("\n".join("%-22s %-20s %.4g" % r for r in report))
    assert not bad, "first failing layers: %s" % bad[:4]
    for a, b in zip(got, want):
        assert mt.heads_rel_err(a, b.numpy()) <= HEAD_TOL


@pytest.mark.parametrize("cfg", [((416, 416, 3), 80, None, 1), ((512, 512, 1), 1, None, 2),
                                 ((512, 512, 1), 1, [(64, 384), (384, 64)], 1), ((608, 608, 3), 80, None, 1)])
def test_heads_vs_oracle(cfg):
    img_size, nc, anchors, B = cfg
    eng, ora = make(img_size, nc, anchors, max_batch=B)
    x = torch.randn(B, img_size[2], img_size[0], img_size[1], generator=torch.Generator().manual_seed(2))
    want = ora.feature_maps(x)
    got = eng.forward_heads(x.numpy())
    errs = [mt.heads_rel_err(a, b.numpy()) for a, b in zip(got, want)]
    print(cfg, errs)
    assert all(a.shape == tuple(b.shape) for a, b in zip(got, want))
    assert max(errs) <= HEAD_TOL, errs


def test_decode_matches_oracle_on_same_heads():
    """decode kernel alone: oracle decode applied to the GPU's own heads (fp32, tolerance 1e-5)"""
    eng, ora = make((416, 416, 3), 80, max_batch=1)
    x = torch.randn(1, 3, 416, 416, generator=torch.Generator().manual_seed(3))
    heads = eng.forward_heads(x.numpy())
    boxes = eng.forward_boxes(x.numpy())
    want = ora.decode([torch.from_numpy(h) for h in heads]).numpy()
    assert boxes.shape == want.shape == (1, 10647, 85)
    np.testing.assert_allclose(boxes, want, rtol=2e-5, atol=2e-4)


def match_fraction(a, b, iou_min=0.99):
    """fraction of boxes in b that have a box in a with IoU >= iou_min (and equal label)"""
    if len(b) == 0:
        return 1.0
    hit = 0
    for row in b:
        cand = a[a[:, 5] == row[5]]
        if len(cand) and np.max(pp.iou_one_vs_many(row[:4].astype(np.float32), cand[:, :4].astype(np.float32))) >= iou_min:
            hit += 1
    return hit / len(b)


def test_detect_end_to_end_sparse_regime():
    """forward -> decode -> filter -> NMS on the device vs oracle forward + reference-pinned C NMS:
    >= 99 % of boxes matched at IoU >= 0.99 (north_star), sparse regime via an objectness bias."""
    img_size, nc = (416, 416, 3), 4
    eng, ora = make(img_size, nc, max_batch=2, obj_bias=-3.0, head_gain=6.0)
    x = torch.randn(2, 3, 416, 416, generator=torch.Generator().manual_seed(4))
    b, s, l, im = eng.detect(x.numpy(), 32, 0.3, 0.1)
    dets = ora(x)
    for i in range(2):
        d = pp.drop_small(dets[i], 32)
        rb, rs, rl = nms_c.class_wise_nms(d[:, :4], d[:, 4:5], d[:, 5:], 0.3, 0.1)
        want = np.concatenate([rb, rs[:, None], rl[:, None].astype(np.float32)], 1)
        m = im == i
        got = np.concatenate([b[m], s[m][:, None], l[m][:, None].astype(np.float32)], 1)
        f1, f2 = match_fraction(got, want), match_fraction(want, got)
        print("image", i, "ref boxes", len(want), "gpu boxes", len(got), "matched", f1, f2)
        assert len(want) > 20
        assert f1 >= 0.99 and f2 >= 0.99
