"""Mint tests/golden/*.npz by running the REFERENCE'S OWN NumPy code verbatim
(imported from /root/reference through oracle/ref_loader.py).  Build-container only.

    python -m oracle.gen_golden            # small fixtures (seconds)
    python -m oracle.gen_golden --full     # + the 200k-box K3 run (~75 s per class)

The fixtures are what pins oracle/postproc_np.py, oracle/tiling_np.py, oracle/nms_oracle.c
(and through them the CUDA path) to the reference.
"""
import argparse
import hashlib
import os
import sys
import time

import numpy as np

from . import cases, ref_loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    args = ap.parse_args()
    ref = ref_loader.load()
    bu, it, ir = ref["bbox_utils"], ref["inference_tiled"], ref["imagereader"]
    os.makedirs(OUT, exist_ok=True)

    # ---- a15/a16: single_class_nms / compute_iou -------------------------------------------
    g = {}
    for tag, (n, canvas, seed, thr) in dict(tiny=(64, 120, 1, 0.3), small=(1500, 700, 2, 0.45),
                                            mid=(12000, 1600, 3, 0.45), loose=(3000, 500, 4, 0.3)).items():
        b, s = cases.nms_case(n, canvas, seed)
        keep = np.asarray(bu.single_class_nms(b, s, thr), dtype=np.int32)
        g[tag + "_boxes"], g[tag + "_scores"], g[tag + "_thr"], g[tag + "_keep"] = b, s, np.float64(thr), keep
    b, s = cases.degenerate_case()
    with np.errstate(all="ignore"):
        g["degen_boxes"], g["degen_scores"], g["degen_thr"] = b, s, np.float64(0.3)
        g["degen_keep"] = np.asarray(bu.single_class_nms(b, s, 0.3), dtype=np.int32)
    b, _ = cases.nms_case(40, 100, 9)
    g["iou_boxes"] = b
    g["iou_row0"] = bu.compute_iou(b[0], b[1:]).astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "nms_single.npz"), **g)

    # ---- a13/a14: filter_small_boxes + per_class_nms ---------------------------------------
    g = {}
    rng = np.random.default_rng(21)
    n, nc = 2500, 6
    bx = cases.boxes_on_canvas(n, 500, 10, 160, rng)
    obj = rng.uniform(0, 1, (n, 1)).astype(np.float32)
    cls = cases.pp.make_tie_free_scores(n * nc, rng, 0.001, 0.999).reshape(n, nc)
    det = np.concatenate([bx, obj, cls], 1).astype(np.float32)
    f = bu.filter_small_boxes(det, 32)
    g["det"], g["filtered_rows"] = det, np.int64(f.shape[0])
    g["filtered_sha"] = np.array(sha(f))
    pb, ps, pl = bu.per_class_nms(f[:, 0:4], f[:, 4:5], f[:, 5:])
    g["pc_boxes"], g["pc_scores"], g["pc_labels"] = pb, ps, pl
    pb2, ps2, pl2 = bu.per_class_nms(f[:, 0:4], f[:, 4:5], f[:, 5:], iou_threshold=0.45, score_threshold=0.6)
    g["pc45_boxes"], g["pc45_scores"], g["pc45_labels"] = pb2, ps2, pl2
    e = bu.per_class_nms(f[:5, 0:4], f[:5, 4:5] * 0, f[:5, 5:])
    assert e == (None, None, None)
    bm, om, cm = cases.multiclass_case(6000, 80, 900, seed=31)
    mb, ms, ml = bu.per_class_nms(bm, om, cm, 0.45, 0.1)
    g["mc80_sha_boxes"], g["mc80_sha_scores"], g["mc80_sha_labels"] = np.array(sha(mb)), np.array(sha(ms)), np.array(sha(ml))
    g["mc80_k"] = np.int64(mb.shape[0])
    np.savez_compressed(os.path.join(OUT, "nms_per_class.npz"), **g)

    # ---- a11/a12: zscore_normalize + convert_image_to_tiles --------------------------------
    g = {}
    for tag, (h, w, c, dt, tile, edge) in dict(
            u16=(700, 900, 1, np.uint16, (512, 512), 96),
            u8rgb=(520, 1100, 3, np.uint8, (256, 320), 64),
            small=(300, 280, 1, np.uint16, (512, 512), 96),       # tile >= image on both axes: r = 0
            wide=(400, 1500, 1, np.uint16, (512, 512), 96)).items():  # r=0 on y only
        img = cases.synthetic_image(h, w, c, dt, seed=hash(tag) % 1000 if False else len(tag) * 7)
        it.EDGE_EFFECT_RANGE = edge
        tiles, xs, ys = it.convert_image_to_tiles(img, list(tile))
        g[tag + "_xs"], g[tag + "_ys"] = np.asarray(xs, np.int64), np.asarray(ys, np.int64)
        g[tag + "_tile_sha"] = np.array([sha(t) for t in tiles])
        g[tag + "_tile_shape"] = np.asarray(tiles[0].shape, np.int64)
        zs = [ir.zscore_normalize(t.astype(np.float32)) for t in tiles]
        g[tag + "_z_mean_std"] = np.asarray([[float(np.mean(t.astype(np.float32))), float(np.std(t.astype(np.float32)))]
                                            for t in tiles], np.float64)
        g[tag + "_z_probe"] = np.asarray([z[5::97, 3::89, 0].ravel()[:16] for z in zs], np.float32)
    flat = np.full((64, 64, 1), 7, np.uint16)
    flat[0, 0, 0] = 8
    g["flat_z"] = ir.zscore_normalize(flat)[:2, :2, 0]            # std <= 1 branch
    it.EDGE_EFFECT_RANGE = 96
    np.savez_compressed(os.path.join(OUT, "tiling.npz"), **g)

    # ---- a17: inference_image_tiled with an injected detector ------------------------------
    g = {}
    for tag, (h, w, c, dt, tile, edge, nb, nc, minbox) in dict(
            e96=(1200, 1500, 1, np.uint16, (512, 512), 96, 900, 2, 32),
            e64=(1000, 1300, 1, np.uint16, (512, 512), 64, 700, 1, 32),
            rgb=(700, 640, 3, np.uint8, (256, 256), 32, 400, 3, 24),
            one=(300, 280, 1, np.uint16, (512, 512), 96, 300, 1, 32)).items():
        img = cases.synthetic_image(h, w, c, dt, seed=100 + len(tag) + edge)
        fake = cases.FakeDetector(nb, nc, tile, seed=5 + edge)
        it.EDGE_EFFECT_RANGE = edge
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            pred = it.inference_image_tiled(fake, img, list(tile), minbox)
        g[tag + "_pred"] = pred
        print("tiled", tag, pred.shape, pred.dtype)
    it.EDGE_EFFECT_RANGE = 96
    np.savez_compressed(os.path.join(OUT, "tiled_pipeline.npz"), **g)

    # ---- union_all_overlapping_bb (off the hot path; bbox_utils.py:138-197) ------------------
    g = {}
    for tag, (n, canvas, seed, thr) in dict(sparse=(60, 600, 1, 0), dense=(120, 300, 2, 0), thr=(150, 400, 3, 0.2),
                                            single=(1, 50, 4, 0)).items():
        b, s = cases.merge_case(n, canvas, seed)
        mb, ms = bu.union_all_overlapping_bb(b.copy(), s.copy(), thr)
        g[tag + "_boxes"] = np.asarray(mb, np.float64)
        g[tag + "_scores"] = np.asarray(ms, np.float64)
        print("merge", tag, n, "->", len(ms))
    np.savez_compressed(os.path.join(OUT, "box_merge.npz"), **g)

    # ---- K3: 200k boxes (reference verbatim, ~75 s) ----------------------------------------
    p = os.path.join(OUT, "nms_k3.npz")
    if args.full or not os.path.exists(p):
        if args.full:
            b, s = cases.k3_single_class()
            t = time.time()
            keep = np.asarray(bu.single_class_nms(b, s, 0.45), dtype=np.int32)
            dt = time.time() - t
            print("K3 200k: kept", keep.size, "in %.1f s" % dt)
            np.savez_compressed(p, keep_sha=np.array(sha(keep)), n_keep=np.int64(keep.size),
                                keep_head=keep[:64], ref_seconds=np.float64(dt), in_sha=np.array(sha(b) + sha(s)))
    for fn in sorted(os.listdir(OUT)):
        print(fn, os.path.getsize(os.path.join(OUT, fn)))


if __name__ == "__main__":
    sys.exit(main())
