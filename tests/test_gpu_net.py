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
    from yolo3_b200 import Y3Error
    for name, ref in ora.trace.items():
        try:
            out = eng.debug_layer_output(name, 2)
        except Y3Error:
            assert name.startswith("conv2d_transpose")      # composed into its consumer conv: no tensor of its own
            continue
        err = mt.heads_rel_err(out, ref.numpy())
        report.append((name, tuple(ref.shape), err))
    bad = [r for r in report if not r[2] < 3e-2]
    print("\n".join("%-22s %-20s %.4g" % r for r in report))
    assert not bad, "first failing layers: %s" % bad[:4]
    for a, b in zip(got, want):
        assert mt.heads_rel_err(a, b.numpy()) <= HEAD_TOL


REF_ANCHORS = [(64, 384), (384, 64)]      # the reference trainer's own anchors (train.py:33)


@pytest.mark.parametrize("cfg", [((416, 416, 3), 80, None, 1), ((512, 512, 1), 1, None, 2),
                                 ((512, 512, 1), 1, REF_ANCHORS, 1),
                                 ((608, 608, 3), 80, None, 1),
                                 # 1-channel, not a multiple of 256 wide, odd batch: partial row segments and an idle image slot
                                 # in the fused stem + conv2d_1 kernel and the halo kernels
                                 ((352, 608, 1), 2, None, 3)])
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


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5])
def test_heads_reference_training_config_over_weight_draws(seed):
    """The reference's own configuration (512x512x1, 1 class, anchors (64,384),(384,64): train.py:33, model.py:108-120)
    over several weight draws.  With bf16 storage of the tail after the first upsample this config sat AT the 2e-2 bar
    (fm3 0.7 / 2.0 / 2.2 % over three draws, fm2 up to 1.8 %); the fp16 tail (net.cu) holds it with a wide margin,
    so the bound asserted here is 1e-2 - a regression of the tail numerics fails the test."""
    eng, ora = make((512, 512, 1), 1, REF_ANCHORS, max_batch=1, seed=seed)
    x = torch.randn(1, 1, 512, 512, generator=torch.Generator().manual_seed(2))
    want = ora.feature_maps(x)
    got = eng.forward_heads(x.numpy())
    errs = [mt.heads_rel_err(a, b.numpy()) for a, b in zip(got, want)]
    print("seed", seed, errs)
    assert max(errs) <= 1e-2, errs


def test_heads_with_random_transposed_conv_weights():
    """upsample_2x is a general Conv2DTranspose(k2,s2) (model.py:94-105): its kernel starts as all ones but is a
    trainable variable.  Random kernel and bias: the composed up-conv path must follow the oracle's conv_transpose."""
    from yolo3_b200 import Engine
    img_size, nc = (256, 320, 3), 3
    W = mt.init_weights(3, nc, 3, seed=7, randomize_bn=True)
    g = torch.Generator().manual_seed(11)
    for name in ("conv2d_transpose", "conv2d_transpose_1"):
        k = W[name + "/kernel"]
        W[name + "/kernel"] = (torch.rand(k.shape, generator=g) * 2 - 1) * (3.0 / k.shape[3]) ** 0.5
        W[name + "/bias"] = torch.randn(k.shape[2], generator=g) * 0.1
    eng = Engine(img_size, nc, None, max_batch=2)
    eng.load_weights({k: v.numpy() for k, v in W.items()})
    ora = mt.OracleNet(W, img_size, nc)
    x = torch.randn(2, 3, 256, 320, generator=torch.Generator().manual_seed(5))
    want = ora.feature_maps(x)
    got = eng.forward_heads(x.numpy())
    errs = [mt.heads_rel_err(a, b.numpy()) for a, b in zip(got, want)]
    print(errs)
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


def calibrated(img_size, nc, anchors, std, obj_bias, x, max_batch, seed=0):
    """random-init weights whose three detection layers are rescaled to logits of the given std and an
    objectness bias (random-init heads differ by 400x in scale: the all-ones upsample inflates them)"""
    from yolo3_b200 import Engine
    A = len(anchors or mt.DEFAULT_ANCHORS)
    W = mt.init_weights(img_size[2], nc, A, seed=seed, randomize_bn=True)
    heads = mt.OracleNet(W, img_size, nc, anchors).feature_maps(x)
    for i, h in enumerate(heads):
        k = "feature_map_%d" % (i + 1)
        W[k + "/kernel"] = W[k + "/kernel"] * (std / float(h.std()))
        b = torch.zeros(A, 5 + nc)
        b[:, 4] = obj_bias
        W[k + "/bias"] = b.reshape(-1)
    eng = Engine(img_size, nc, anchors, max_batch=max_batch)
    eng.load_weights({k: v.numpy() for k, v in W.items()})
    return eng, mt.OracleNet(W, img_size, nc, anchors)


def test_detect_pipeline_exact_on_own_boxes():
    """decode -> filter_small -> per_class_nms on the device == the reference-pinned oracle applied to the
    SAME decoded boxes (the GPU's own forward_boxes output): bit-exact boxes, scores, labels, order -
    in a dense, chaotic NMS regime (default anchors, every box a candidate)."""
    x = torch.randn(2, 3, 416, 416, generator=torch.Generator().manual_seed(4))
    eng, _ = calibrated((416, 416, 3), 4, None, 0.25, -3.0, x[:1], 2)
    dec = eng.forward_boxes(x.numpy())
    b, s, l, im = eng.detect(x.numpy(), 32, 0.3, 0.1)
    for i in range(2):
        d = pp.drop_small(dec[i], 32)
        rb, rs, rl = nms_c.class_wise_nms(d[:, :4], d[:, 4:5], d[:, 5:], 0.3, 0.1)
        m = im == i
        assert len(rb) > 100
        assert np.array_equal(b[m], rb) and np.array_equal(s[m], rs) and np.array_equal(l[m], rl)


def test_detect_end_to_end_separated_regime():
    """north_star: end-to-end detection sets match the fp32 oracle at IoU >= 0.99 for >= 99 % of boxes.
    Checked in a SEPARATED regime (one 12x12 anchor: boxes smaller than the 8-px grid pitch of the finest
    scale overlap below the NMS threshold), where greedy NMS has no long near-tie suppression chains;
    with the default 32..256-px anchors every cell's boxes overlap dozens of neighbours with
    near-equal scores and a 1e-3 score perturbation legitimately reorders the greedy picks
    (the boxes themselves still agree: see DESIGN.md, "end-to-end criterion")."""
    anchors = [(12, 12)]
    x = torch.randn(2, 3, 416, 416, generator=torch.Generator().manual_seed(4))
    eng, ora = calibrated((416, 416, 3), 2, anchors, 0.2, -3.0, x[:1], 2)
    b, s, l, im = eng.detect(x.numpy(), 0, 0.3, 0.1)
    dets = ora(x)
    for i in range(2):
        d = pp.drop_small(dets[i], 0)
        rb, rs, rl = nms_c.class_wise_nms(d[:, :4], d[:, 4:5], d[:, 5:], 0.3, 0.1)
        want = np.concatenate([rb, rs[:, None], rl[:, None].astype(np.float32)], 1)
        m = im == i
        got = np.concatenate([b[m], s[m][:, None], l[m][:, None].astype(np.float32)], 1)
        f1, f2 = match_fraction(got, want), match_fraction(want, got)
        print("image", i, "ref boxes", len(want), "gpu boxes", len(got), "matched", f1, f2)
        assert len(want) > 1000
        assert f1 >= 0.99 and f2 >= 0.99


def test_forward_is_deterministic_at_bench_batch():
    """races in the warp-specialised kernels show up as run-to-run differences (or launch failures) once every CTA
    walks many tiles: 48 tiles of 512x512x1 (fused stem + conv2d_1, halo kernels, 2-CTA kernels), three runs"""
    from yolo3_b200 import Engine, weights
    B = 48
    eng = Engine((512, 512, 1), 1, None, max_batch=B)
    eng.load_weights(weights.random_init(1, 1, 3, seed=0, randomize_bn=True))
    x = np.random.default_rng(0).standard_normal((B, 1, 512, 512)).astype(np.float32)
    first = [h.copy() for h in eng.forward_heads(x)]
    assert all(np.isfinite(h).all() for h in first)
    for _ in range(2):
        again = eng.forward_heads(x)
        assert all(np.array_equal(a, b) for a, b in zip(first, again))
    # the same images in a smaller batch give the same heads (tile -> CTA assignment must not matter)
    part = eng.forward_heads(x[:5])
    assert all(np.array_equal(a[:5], b) for a, b in zip(first, part))


def test_detect_image_is_the_reference_flow_in_one_call():
    """y3_detect_image (inference.py:47-79 in one library call, CUDA-graph forward): whole-image z-score, forward, decode,
    clip to the image, small-box filter, per-class NMS == the same stages run one by one (oracle z-score, the GPU's own
    decoded boxes, then the reference-pinned NumPy / C post-processing) - bit-exact, twice (eager first call, graph replay)."""
    from oracle import tiling_np as tl
    img = (np.random.default_rng(9).integers(0, 4000, (416, 416, 3))).astype(np.uint16)
    xn = tl.zscore(img.astype(np.float32)).transpose(2, 0, 1)[None]
    x1 = torch.from_numpy(np.ascontiguousarray(xn))
    eng, _ = calibrated((416, 416, 3), 4, None, 0.25, -3.0, x1, 2)
    # the GPU's own z-score of the image (one tile = the whole image; within 2e-6 of the oracle's fp32 z-score, which is
    # pinned separately in test_gpu_tiles) so that the stages after it can be compared bit for bit
    xg = eng.tiles_normalized(img, (416, 416), 96, 0, 1)
    np.testing.assert_allclose(xg, xn, rtol=0, atol=1e-5)
    dec = eng.forward_boxes(xg)[0].copy()
    for col, hi in ((0, 416), (1, 416), (2, 416), (3, 416)):
        dec[:, col] = np.clip(dec[:, col], 0, hi)
    d = pp.drop_small(dec, 32)
    rb, rs, rl = nms_c.class_wise_nms(d[:, :4], d[:, 4:5], d[:, 5:], 0.3, 0.1)
    assert len(rb) > 100 and (rb.min() == 0 or rb.max() == 416)          # the clip really acts on this input
    for _ in range(3):
        b, s, l = eng.detect_image(img, 32, 0.3, 0.1)
        assert np.array_equal(b, rb) and np.array_equal(s, rs) and np.array_equal(l, rl)
    # without the clip: the plain detect() of the normalised image
    b0, s0, l0, _ = eng.detect(xg, 32, 0.3, 0.1)
    b1, s1, l1 = eng.detect_image(img, 32, 0.3, 0.1, clip=False)
    assert np.array_equal(b0, b1) and np.array_equal(s0, s1) and np.array_equal(l0, l1)


def test_graph_replay_equals_eager_forward():
    """small batches replay a captured CUDA graph of the forward: same heads as the eager launches, also after a weight reload"""
    eng, ora = make((256, 256, 1), 1, max_batch=8)
    x = np.random.default_rng(3).standard_normal((8, 1, 256, 256)).astype(np.float32)
    big = eng.forward_heads(x)                      # batch 8: eager
    for b in (1, 2, 3):
        first = [h.copy() for h in eng.forward_heads(x[:b])]            # eager + capture
        again = eng.forward_heads(x[:b])                                 # replay
        assert all(np.array_equal(p, q) for p, q in zip(first, again))
        assert all(np.array_equal(p, q[:b]) for p, q in zip(first, big))
    W2 = mt.init_weights(1, 1, 3, seed=5, randomize_bn=True)
    eng.load_weights({k: v.numpy() for k, v in W2.items()})
    got = eng.forward_heads(x[:1])
    want = mt.OracleNet(W2, (256, 256, 1), 1).feature_maps(torch.from_numpy(x[:1]))
    assert max(mt.heads_rel_err(a, b.numpy()) for a, b in zip(got, want)) <= HEAD_TOL


def test_results_do_not_depend_on_stale_activations():
    """Own bounds check (compute-sanitizer is closed on this pool): poison every activation buffer and image slot with a
    full batch of 50x larger inputs, then run smaller batches - the heads must equal a fresh handle's, bit for bit.  A tile
    that read rows of a neighbouring (stale) image slot, or a partial M tile leaking into valid rows, would show."""
    from yolo3_b200 import Engine, weights
    w = weights.random_init(1, 1, 3, seed=0, randomize_bn=True)
    x = np.random.default_rng(0).standard_normal((6, 1, 512, 512)).astype(np.float32)
    fresh = Engine((512, 512, 1), 1, None, max_batch=6)
    fresh.load_weights(w)
    want = {b: [h.copy() for h in fresh.forward_heads(x[:b])] for b in (1, 3, 5)}
    fresh.close()
    eng = Engine((512, 512, 1), 1, None, max_batch=6)
    eng.load_weights(w)
    for b in (1, 3, 5):
        eng.forward_heads(x * 50.0)
        got = eng.forward_heads(x[:b])
        assert all(np.array_equal(p, q) for p, q in zip(got, want[b])), "batch %d depends on stale buffers" % b
