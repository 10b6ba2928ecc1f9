"""The drop-in scripts (object-detection-yolov3_b200/inference.py, inference_tiled.py) end to end on the GPU:
same functions, arguments and command lines as the reference (inference.py:24-135, inference_tiled.py:313-382),
CSV outputs checked against direct library calls + the reference-pinned oracle."""
import csv
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import cases, nms_c, postproc_np as pp, tiling_np as tl

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "object-detection-yolov3_b200")


def make_model_dir(path, img_size, nc, seed=0):
    from yolo3_b200 import weights
    w = weights.random_init(img_size[2], nc, 3, seed=seed, randomize_bn=True)
    for i in (1, 2, 3):                                   # tame, sparse heads (see bench.calibrate_heads)
        k = "feature_map_%d" % i
        w[k + "/kernel"] = (w[k + "/kernel"] * [2.0, 0.05, 0.004][i - 1]).astype(np.float32)
        b = np.zeros((3, 5 + nc), np.float32)
        b[:, 4] = -2.0
        w[k + "/bias"] = b.reshape(-1)
    weights.save_model_dir(path, w, img_size, nc, [(32, 32), (64, 64), (128, 128)])
    return w


def read_csv(p):
    with open(p) as fh:
        rows = list(csv.reader(fh))
    return rows[0], rows[1:]


def test_inference_tiled_folder_and_cli(tmp_path):
    import cv2
    import inference_tiled
    import model
    model_dir, img_dir, out_dir = tmp_path / "model", tmp_path / "images", tmp_path / "out"
    img_dir.mkdir()
    make_model_dir(str(model_dir), (256, 256, 1), 2)
    imgs = {}
    for name, (h, w) in dict(a=(700, 900), b=(300, 520)).items():
        img = cases.synthetic_image(h, w, 1, np.uint16, seed=h, blobs=15)
        assert cv2.imwrite(str(img_dir / (name + ".tif")), img[:, :, 0])
        imgs[name] = img
    (img_dir / "ignored.png").write_bytes(b"x")
    inference_tiled.inference_image_folder(str(img_dir), "tif", str(model_dir), str(out_dir), [256, 256], 24)
    m = model.load_saved_model(str(model_dir), max_batch=8)
    for name, img in imgs.items():
        hdr, rows = read_csv(out_dir / (name + ".csv"))
        assert hdr == ["X", "Y", "W", "H", "P", "C"]
        pred = m.engine.infer_tiled(img, (256, 256), 24, inference_tiled.EDGE_EFFECT_RANGE)
        assert len(rows) == pred.shape[0] and len(rows) > 0
        for r, p in zip(rows, pred):
            x, y = int(p[0]), int(p[1])
            assert [int(r[0]), int(r[1]), int(r[2]), int(r[3]), int(r[5])] == [x, y, int(p[2] - x + 1), int(p[3] - y + 1), int(p[5])]
            assert r[4] == "%f" % p[4]
        # and the library result is the reference pipeline applied to the library's own decoded boxes
        tiles = m.engine.tiles_normalized(img, (256, 256), inference_tiled.EDGE_EFFECT_RANGE)
        dets = np.concatenate([m.engine.forward_boxes(tiles[i:i + 8]) for i in range(0, len(tiles), 8)])
        it = iter(range(len(tiles)))
        want = tl.tiled_inference(lambda x: dets[next(it)][None], img, (256, 256), 24, edge_range=96, nms_fn=nms_c.greedy_nms)
        assert np.array_equal(pred, want)
    # same thing through the unchanged command line
    out2 = tmp_path / "out_cli"
    env = dict(os.environ, PYTHONPATH=PKG)
    subprocess.check_call([sys.executable, os.path.join(PKG, "inference_tiled.py"), "--saved-model-filepath", str(model_dir),
                           "--image-folder", str(img_dir), "--output-folder", str(out2), "--tile-height", "256",
                           "--tile-width", "256", "--min-box-size", "24", "--image-format", "tif"], env=env, cwd=PKG,
                          stdout=subprocess.DEVNULL)
    for name in imgs:
        assert (out2 / (name + ".csv")).read_text() == (out_dir / (name + ".csv")).read_text()
    # a foreign model callable goes through the generic path (GPU slicing + GPU stitching)
    fake = cases.FakeDetector(300, 2, (256, 256), seed=3)
    got = inference_tiled.inference_image_tiled(fake, imgs["a"], [256, 256], 24)
    gpu_tiles = m.engine.tiles_normalized(imgs["a"], (256, 256), 96)      # the detector sees the GPU-normalised tiles
    it = iter(range(len(gpu_tiles)))
    want = tl.tiled_inference(lambda x: fake(gpu_tiles[next(it)][None]), imgs["a"], (256, 256), 24, edge_range=96)
    assert np.array_equal(got, want)
    # convert_image_to_tiles keeps the reference's return convention
    t, xs, ys = inference_tiled.convert_image_to_tiles(imgs["b"], [256, 256])
    t2, xs2, ys2 = tl.cut_tiles(imgs["b"], (256, 256), 96)
    assert xs == xs2 and ys == ys2 and all(np.array_equal(a, b) for a, b in zip(t, t2))


def test_inference_single_image_folder_and_cli(tmp_path):
    import cv2
    import bbox_utils
    import imagereader
    import inference
    import model
    model_dir, img_dir, out_dir = tmp_path / "model", tmp_path / "images", tmp_path / "out"
    img_dir.mkdir()
    make_model_dir(str(model_dir), (256, 256, 3), 2)
    img = cases.synthetic_image(256, 256, 3, np.uint8, seed=9, blobs=6)
    assert cv2.imwrite(str(img_dir / "one.png"), img[:, :, ::-1])
    inference.inference(str(img_dir), ".png", str(model_dir), str(out_dir), 24)
    hdr, rows = read_csv(out_dir / "one.csv")
    assert hdr == ["X", "Y", "W", "H", "C"]
    # reproduce with library calls + the oracle NMS
    m = model.load_saved_model(str(model_dir))
    z = imagereader.zscore_normalize(imagereader.imread(str(img_dir / "one.png")).astype(np.float32))
    np.testing.assert_allclose(z, tl.zscore(img.astype(np.float32)), rtol=2e-6, atol=2e-6)
    boxes = np.array(m(z.transpose(2, 0, 1)[None], training=False))[0]
    for c, hi in ((0, 256), (1, 256), (2, 256), (3, 256)):
        boxes[:, c] = np.clip(boxes[:, c], 0, hi)
    f = pp.drop_small(boxes, 24)
    b, s, l = nms_c.class_wise_nms(f[:, :4], f[:, 4:5], f[:, 5:])
    b[:, 2] -= b[:, 0]
    b[:, 3] -= b[:, 1]
    want = np.concatenate((b, l.reshape(-1, 1)), axis=-1).astype(np.int32)
    assert len(rows) == len(want) > 0
    assert [[int(v) for v in r] for r in rows] == want.tolist()
    out2 = tmp_path / "out_cli"
    subprocess.check_call([sys.executable, os.path.join(PKG, "inference.py"), "--saved-model-filepath", str(model_dir),
                           "--image-folder", str(img_dir), "--output-folder", str(out2), "--image-format", "png",
                           "--min-box-size", "24"], env=dict(os.environ, PYTHONPATH=PKG), cwd=PKG, stdout=subprocess.DEVNULL)
    assert (out2 / "one.csv").read_text() == (out_dir / "one.csv").read_text()
    # bbox_utils facade == oracle on the same rows
    assert bbox_utils.single_class_nms(f[:200, :4], f[:200, 4], 0.3) == nms_c.greedy_nms(f[:200, :4], f[:200, 4], 0.3)


def test_tf_saved_model_directory_is_a_drop_in(tmp_path):
    """--saved-model-filepath pointing at the reference's own format (saved_model.pb + variables/ bundle, train.py:221),
    read without TensorFlow, gives the same CSV as the side-car format holding the same variables."""
    import cv2
    import inference_tiled
    from yolo3_b200 import tf_bundle
    side_dir, tf_dir, img_dir = tmp_path / "side", tmp_path / "saved_model", tmp_path / "images"
    img_dir.mkdir()
    w = make_model_dir(str(side_dir), (256, 256, 1), 2)
    tf_bundle.write_saved_model_variables(str(tf_dir), w, input_shape=[-1, 1, 256, 256], checksum_limit=1 << 16)
    import json
    json.dump({"anchors": [[32, 32], [64, 64], [128, 128]]}, open(str(tf_dir / "y3_config.json"), "w"))
    img = cases.synthetic_image(600, 520, 1, np.uint16, seed=5, blobs=12)
    assert cv2.imwrite(str(img_dir / "a.tif"), img[:, :, 0])
    inference_tiled.inference_image_folder(str(img_dir), "tif", str(side_dir), str(tmp_path / "o1"), [256, 256], 24)
    inference_tiled.inference_image_folder(str(img_dir), "tif", str(tf_dir), str(tmp_path / "o2"), [256, 256], 24)
    a, b = (tmp_path / "o1" / "a.csv").read_text(), (tmp_path / "o2" / "a.csv").read_text()
    assert a == b and a.count("\n") > 3
    # a model directory that does not pin the input size serves any tile size (the network is fully convolutional)
    os.remove(str(tf_dir / "saved_model.pb"))
    inference_tiled.inference_image_folder(str(img_dir), "tif", str(tf_dir), str(tmp_path / "o3"), [256, 256], 24)
    assert (tmp_path / "o3" / "a.csv").read_text() == a
    # the optional cross-seam stage can only remove rows
    inference_tiled.CROSS_SEAM_NMS = True
    try:
        inference_tiled.inference_image_folder(str(img_dir), "tif", str(side_dir), str(tmp_path / "o4"), [256, 256], 24)
    finally:
        inference_tiled.CROSS_SEAM_NMS = False
    rows4 = (tmp_path / "o4" / "a.csv").read_text().splitlines()
    assert set(rows4) <= set(a.splitlines()) and len(rows4) >= 2


def test_bbox_utils_off_path_helpers(tmp_path, golden):
    """the non-hot-path helpers of the bbox_utils facade: box merge (IoU on the GPU) against the reference's output,
    CSV writers / loaders round trip"""
    import bbox_utils
    g = golden("box_merge.npz")
    for tag, (n, canvas, seed, thr) in dict(sparse=(60, 600, 1, 0), dense=(120, 300, 2, 0), thr=(150, 400, 3, 0.2),
                                            single=(1, 50, 4, 0)).items():
        b, s = cases.merge_case(n, canvas, seed)
        mb, ms = bbox_utils.union_all_overlapping_bb(b.copy(), s.copy(), thr)
        assert np.array_equal(np.asarray(mb, np.float64), g[tag + "_boxes"]), tag
        assert np.allclose(np.asarray(ms, np.float64), g[tag + "_scores"], rtol=0, atol=1e-12), tag
    rows = np.array([[5, 7, 20, 31, 1], [0, 0, 3, 3, 0]], np.int64)
    p = str(tmp_path / "b.csv")
    bbox_utils.write_boxes_from_ltrbc(rows, p)
    assert np.array_equal(bbox_utils.load_boxes_to_ltrbc(p), rows.astype(np.float64))
    assert bbox_utils.load_boxes_to_xywhc(p).tolist() == [[5, 7, 16, 25, 1], [0, 0, 4, 4, 0]]
    assert bbox_utils.load_boxes_to_ltrbc(str(tmp_path / "missing.csv")).shape == (0, 5)
