"""bench.py - headline benchmark of the tiled YOLOv3 inference hot path (BASELINE.json metric:
image megapixels/s of inference_tiled on a synthetic 20000x20000 uint16 image, 512x512 tiles,
64-px overlap, 1/2/4/8 B200).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

A step = one pass of the whole hot path over the whole image: tile slicing + per-tile z-score ->
Darknet-53 + 3 heads (tcgen05) -> decode -> small-box filter + per-class NMS -> ownership stitch.
Under torchrun the tile grid is sharded across ranks (strong scaling: the image is fixed) and the
result boxes are all-gathered with NCCL.

`value`  : image Mpix/s with the image already resident in HBM.
`e2e`    : the same through the public API with the image in pinned HOST memory: the H2D copy of
           the rank's row band and the D2H read of the result boxes are inside the timed region.
TensorFlow is not installed (and cannot be), so the reference's TF path cannot be timed anywhere;
the CPU baseline is the oracle port: torch-CPU fp32 restatement of model.py + the NumPy
restatement of the reference's post-processing (pinned to the reference's own code by tests).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "object-detection-yolov3_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

TILE = (512, 512)
NC, ANCHORS = 1, [(32, 32), (128, 128), (256, 256)]
CONV_GF_PER_TILE = 99.00130304          # sum 2*M*N*K over the 75 Conv2D, 512x512x1, A=3, NC=1
MIN_BOX, IOU_THR, SCORE_THR = 32, 0.3, 0.1


def synthetic_image(side, seed=7, blobs=4000):
    """K4 input: uniform uint16 noise with planted bright blobs."""
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 65535, (side, side, 1), dtype=np.uint16)
    for _ in range(blobs):
        y, x, r = int(rng.integers(0, side)), int(rng.integers(0, side)), int(rng.integers(8, 40))
        img[max(0, y - r):y + r, max(0, x - r):x + r] = 65535
    return img


def bench_weights(seed=0):
    from yolo3_b200 import weights
    return weights.random_init(1, NC, len(ANCHORS), seed=seed, randomize_bn=True)


def calibrate_heads(eng, w, sample_tiles, target_std=1.0, pass_frac=0.005, nc=NC, n_anchors=len(ANCHORS), interior=True):
    """Random-init heads are useless as a detection regime: the all-ones upsample inflates the three heads
    by 400x relative to each other and every output channel is a large constant plus a small spatial
    signal, so whole channels pass or fail the score threshold together.  Standardise every detection
    channel (measured on the GPU path itself over sample tiles) to logits ~ N(0, target_std) and shift
    the objectness channels so that about `pass_frac` of the anchors clear objectness 0.02 (score >= 0.1
    needs obj*cls >= 0.01) -> a sparse, spatially varying set of candidates."""
    heads = eng.forward_heads(sample_tiles)
    E = 5 + nc
    upd = {}
    for i, h in enumerate(heads):
        k = "feature_map_%d" % (i + 1)
        g = h.shape[2]
        m = max(1, g // 4) if interior else 0               # interior cells only: the zero padding at tile
        core = h[:, :, m:g - m, m:g - m]                    # borders would otherwise dominate the statistics
        mu = core.mean(axis=(0, 2, 3)).astype(np.float64)
        sd = np.maximum(core.std(axis=(0, 2, 3)).astype(np.float64), 1e-12)
        tgt = np.full((n_anchors, E), target_std)
        tgt[:, 2:4] = 0.3                                   # box-size logits: boxes stay near their anchor size
        gain = tgt.reshape(-1) / sd
        z = (core - mu[None, :, None, None]) / sd[None, :, None, None]          # standardised logits of the sample
        zobj = z.reshape(z.shape[0], n_anchors, E, -1)[:, :, 4, :]
        obj_bias = -3.9 - float(np.quantile(zobj, 1.0 - pass_frac)) * target_std
        want = np.zeros(n_anchors * E)
        want.reshape(n_anchors, E)[:, 4] = obj_bias
        upd[k + "/kernel"] = (w[k + "/kernel"] * gain[None, None, None, :]).astype(np.float32)
        upd[k + "/bias"] = (want + (w[k + "/bias"] - mu) * gain).astype(np.float32)
    eng.load_weights(upd)
    w.update(upd)


class ClockSampler(threading.Thread):
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [v.strip() for v in out.strip().split(",")]
                if len(f) >= 6:
                    self.rows.append(f)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(r[2 + j].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


def cpu_baseline_run(img, edge, n_tiles_sample, threads):
    """The oracle port on host cores: torch fp32 forward + NumPy post-processing on the first
    `n_tiles_sample` tiles of the same workload.  Returns (Mpix/s of image area, seconds)."""
    import torch
    from oracle import model_torch as mt, tiling_np as tl, postproc_np as pp, nms_c
    torch.set_num_threads(threads)
    w = bench_weights()
    ora = mt.OracleNet({k: torch.from_numpy(v) for k, v in w.items()}, TILE + (1,), NC, ANCHORS)
    plan, (ry, rx) = tl.tile_plan(img.shape[0], img.shape[1], TILE, edge)
    zone = (TILE[0] - 2 * ry) * (TILE[1] - 2 * rx)
    t0 = time.perf_counter()
    for p in plan[:n_tiles_sample]:
        t = img[p["y0"]:p["y1"], p["x0"]:p["x1"]]
        (pt, pb), (pl, pr) = p["pad"]
        if pt or pb or pl or pr:
            t = np.pad(t, ((pt, pb), (pl, pr), (0, 0)), mode="reflect")
        x = tl.zscore(t.astype(np.float32)).transpose(2, 0, 1)[None]
        det = ora(np.ascontiguousarray(x))[0]
        det = pp.drop_small(det, MIN_BOX)
        b, s, l = pp.class_wise_nms(det[:, :4], det[:, 4:5], det[:, 5:], IOU_THR, SCORE_THR, nms_fn=nms_c.greedy_nms)
        if b is not None:
            with np.errstate(invalid="ignore", over="ignore"):          # random-weight boxes can be inf: the reference warns too
                tl.ghost_band_mask(b, p["rec_x"], p["rec_y"], img.shape[:2], TILE, edge)
    dt = time.perf_counter() - t0
    return n_tiles_sample * zone / 1e6 / dt, dt


# ------------------------------------------------------------------------------------------ auxiliary measurements
# The other BASELINE.json configurations, each with a parity flag computed in the same run against the oracle
# (test infrastructure used as the CHECKER only - nothing timed below runs oracle code, except the labelled CPU numbers).
def _head_err(eng, w, img_size, nc, x):
    import torch
    from oracle import model_torch as mt
    ora = mt.OracleNet({k: torch.from_numpy(np.asarray(v)) for k, v in w.items()}, img_size, nc, ANCHORS)
    want = ora.feature_maps(torch.from_numpy(x))
    got = eng.forward_heads(x)
    return [float(mt.heads_rel_err(a, b.numpy())) for a, b in zip(got, want)]


def aux_k3(local, args, peaks, img, w):
    """NMS stress (BASELINE configs[2]): 200 k candidates, IoU 0.45 - 1 class and 80 classes, boxes/s."""
    from oracle import cases, nms_c
    from yolo3_b200 import post_engine
    pe = post_engine(local)
    rng = np.random.default_rng(0)
    c = rng.uniform(0, 2000, (200_000, 2))
    wh = rng.uniform(33, 300, (200_000, 2))
    kb = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
    ks = rng.permutation(200_000).astype(np.float32) / 200_000
    pe.single_class_nms(kb, ks, 0.45)
    t0 = time.perf_counter()
    keep = pe.single_class_nms(kb, ks, 0.45)
    dt_nms = time.perf_counter() - t0
    tn = pe.timings()
    out = {"boxes": 200_000, "classes": 1, "iou_thr": 0.45, "kept": int(keep.size), "boxes_per_s_e2e_host_arrays": 200_000 / dt_nms,
           "boxes_per_s_device": 200_000 / (tn["ms_nms"] * 1e-3), "ms_nms_device": tn["ms_nms"],
           "kept_equals_c_oracle_on_first_20k": bool(pe.single_class_nms(kb[:20000], ks[:20000], 0.45).tolist()
                                                     == nms_c.greedy_nms(kb[:20000], ks[:20000], 0.45))}
    b, o, cm = cases.multiclass_case(200_000, 80, 2000, seed=3)
    pe.per_class_nms(b, o, cm, 0.45, 0.1)
    t0 = time.perf_counter()
    G = pe.per_class_nms(b, o, cm, 0.45, 0.1)
    dt80 = time.perf_counter() - t0
    t80 = pe.timings()
    R = nms_c.class_wise_nms(b, o, cm, 0.45, 0.1)
    out["classes_80"] = {"candidates": int(t80["candidates"]), "kept": int(t80["kept"]), "ms_nms_device": t80["ms_nms"],
                         "candidates_per_s_device": t80["candidates"] / (t80["ms_nms"] * 1e-3),
                         "boxes_per_s_e2e_host_arrays": 200_000 / dt80,
                         "bit_exact_vs_c_oracle": bool(all(np.array_equal(x, y) for x, y in zip(R, G)))}
    return out


def aux_k2(local, args, peaks, img, w):
    """BASELINE configs[1]: 416x416x3, batch 64, NC=80 - conv stack TFLOP/s and decode+NMS against HBM bandwidth on SURVEY
    8(d)'s byte model B*N*(5+NC)*4 (heads read once) + K_cand*56 + k_kept*4."""
    from oracle import nms_c, postproc_np as pp
    from yolo3_b200 import Engine, weights as _wts
    e2 = Engine((416, 416, 3), 80, ANCHORS, max_batch=64, device=local)
    w2 = _wts.random_init(3, 80, 3, seed=0, randomize_bn=True)
    e2.load_weights(w2)
    x2 = np.random.default_rng(1).standard_normal((64, 3, 416, 416)).astype(np.float32)
    calibrate_heads(e2, w2, x2[:8], pass_frac=0.002, nc=80, interior=False)
    best = None
    for _ in range(6):
        rb, rs, rl, ri = e2.detect(x2, MIN_BOX, IOU_THR, SCORE_THR)
        t2 = e2.timings()
        if best is None or t2["ms_nms"] < best["ms_nms"]:
            best = t2
    t2 = best
    nbytes = 64 * 10647 * 85 * 4 + t2["candidates"] * 56 + t2["kept"] * 4
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    # parity: the device pipeline on image 0 == reference-pinned post-processing of the GPU's own decoded boxes
    d0 = pp.drop_small(e2.forward_boxes(x2[:1])[0], MIN_BOX)
    ob, os_, ol = nms_c.class_wise_nms(d0[:, :4], d0[:, 4:5], d0[:, 5:], IOU_THR, SCORE_THR)
    m = ri == 0
    exact = bool(ob is not None and np.array_equal(rb[m], ob) and np.array_equal(rs[m], os_) and np.array_equal(rl[m], ol))
    return {"conv_ms": t2["ms_conv"], "conv_tflops": 64 * 66.12988928e9 / (t2["ms_conv"] * 1e-3) / 1e12,
            "decode_nms_ms": t2["ms_nms"], "candidates": int(t2["candidates"]), "kept": int(t2["kept"]),
            "decode_nms_algorithmic_gbytes_per_s": nbytes / (t2["ms_nms"] * 1e-3) / 1e9,
            "decode_nms_frac_of_hbm": nbytes / (t2["ms_nms"] * 1e-3) / 1e9 / hbm,
            "decode_nms_dram_bytes_note": "the fused path reads one objectness logit per row and full rows only where the row can pass; "
                                          "bytes actually moved are in profiles/r2_ncu_k2_post_*.txt - the fraction on ACTUAL bytes is lower",
            "decode_threshold_compact_kernels_ms": t2["ms_decode"],
            "decode_threshold_compact_frac_of_hbm": 64 * 10647 * 85 * 4 / (t2["ms_decode"] * 1e-3) / 1e9 / hbm,
            "images_per_s_device": 64 / ((t2["ms_conv"] + t2["ms_nms"]) * 1e-3),
            "pipeline_bit_exact_on_own_boxes_image0": exact}


def aux_k1(local, args, peaks, img, w):
    """BASELINE configs[0]: 416x416x3, batch 1, NC=80 through the inference.py-equivalent single call (y3_detect_image:
    z-score + CUDA-graph forward + decode + clip + filter + NMS), host uint8 image in, host boxes out - ms per image,
    next to the oracle port on the host cores."""
    import torch
    from oracle import model_torch as mt, nms_c, postproc_np as pp, tiling_np as tl
    from yolo3_b200 import Engine, weights as _wts
    e1 = Engine((416, 416, 3), 80, ANCHORS, max_batch=1, device=local)
    w1 = _wts.random_init(3, 80, 3, seed=0, randomize_bn=True)
    e1.load_weights(w1)
    im = np.random.default_rng(1234).integers(0, 256, (416, 416, 3), dtype=np.uint8)
    xn = e1.tiles_normalized(im, (416, 416), 96, 0, 1)
    calibrate_heads(e1, w1, xn, pass_frac=0.002, nc=80, interior=False)
    for _ in range(5):
        e1.detect_image(im, MIN_BOX, IOU_THR, SCORE_THR)
    ts, dev_ms = [], []
    for _ in range(30):
        t0 = time.perf_counter()
        b, sc, lb = e1.detect_image(im, MIN_BOX, IOU_THR, SCORE_THR)
        ts.append(1e3 * (time.perf_counter() - t0))
        dev_ms.append(e1.timings())
    tm = dev_ms[int(np.argsort([d["ms_total"] for d in dev_ms])[len(dev_ms) // 2])]
    # parity: same stages one by one on the GPU's own decoded boxes (bit-exact), heads vs the fp32 oracle
    dec = e1.forward_boxes(xn)[0].copy()
    for col in range(4):
        dec[:, col] = np.clip(dec[:, col], 0, 416)
    d = pp.drop_small(dec, MIN_BOX)
    ob, os_, ol = nms_c.class_wise_nms(d[:, :4], d[:, 4:5], d[:, 5:], IOU_THR, SCORE_THR)
    exact = bool((ob is None and b.shape[0] == 0) or (ob is not None and np.array_equal(b, ob) and np.array_equal(sc, os_) and np.array_equal(lb, ol)))
    herr = _head_err(e1, w1, (416, 416, 3), 80, xn)
    # oracle port on the host cores (labelled CPU number): fp32 forward + decode + post-processing of the same image
    ora = mt.OracleNet({k: torch.from_numpy(np.asarray(v)) for k, v in w1.items()}, (416, 416, 3), 80, ANCHORS)
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    xo = tl.zscore(im.astype(np.float32)).transpose(2, 0, 1)[None]
    det = ora(np.ascontiguousarray(xo))[0]
    for col in range(4):
        det[:, col] = np.clip(det[:, col], 0, 416)
    dd = pp.drop_small(det, MIN_BOX)
    pp.class_wise_nms(dd[:, :4], dd[:, 4:5], dd[:, 5:], IOU_THR, SCORE_THR, nms_fn=nms_c.greedy_nms)
    cpu_ms = 1e3 * (time.perf_counter() - t0)
    return {"ms_per_image_wall_median": float(np.median(ts)), "ms_per_image_wall_min": float(np.min(ts)),
            "ms_device_total": tm["ms_total"], "ms_device_h2d": tm["ms_h2d"], "ms_device_zscore": tm["ms_prep"], "ms_device_conv": tm["ms_conv"],
            "ms_device_decode_nms": tm["ms_nms"], "boxes": int(b.shape[0]), "forward": "CUDA graph replay (77 launches)",
            "cpu_port_ms_per_image": cpu_ms, "cpu_cores": os.cpu_count(), "speedup_vs_cpu_port": cpu_ms / float(np.median(ts)),
            "pipeline_bit_exact_on_own_boxes": exact, "heads_rel_err_vs_fp32_oracle": herr, "heads_within_2e-2": bool(max(herr) <= 2e-2)}


def aux_k4_edge96(local, args, peaks, img, w):
    """K4 with the reference's own EDGE_EFFECT_RANGE = 96 (inference_tiled.py:26): 3969 tiles instead of 2809."""
    import torch
    from yolo3_b200 import Engine, tile_count
    side = img.shape[0]
    n_tiles = tile_count(side, side, TILE, 96)
    e4 = Engine(TILE + (1,), NC, ANCHORS, max_batch=256, device=local)
    e4.load_weights(w)                                   # the calibrated bench weights
    dev = torch.device("cuda", local)
    img_dev = torch.from_numpy(img.view(np.int16)).to(dev).view(torch.uint16)
    e4.infer_tiled(img_dev, TILE, MIN_BOX, 96, IOU_THR, SCORE_THR, out_device=dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n_rep = 3
    for _ in range(n_rep):
        out = e4.infer_tiled(img_dev, TILE, MIN_BOX, 96, IOU_THR, SCORE_THR, out_device=dev)
    torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - t0) / n_rep
    t = e4.timings()
    sample = e4.tiles_normalized(np.ascontiguousarray(img[:1024, :2048]), TILE, 96, 0, 2)
    herr = _head_err(e4, w, TILE + (1,), NC, sample)
    return {"edge_range": 96, "tiles": n_tiles, "ms_per_step": ms, "image_mpix_per_s": side * side / 1e6 / (ms * 1e-3),
            "network_input_mpix_per_s": n_tiles * TILE[0] * TILE[1] / 1e6 / (ms * 1e-3), "boxes": int(out.shape[0]),
            "conv_tflops": n_tiles * CONV_GF_PER_TILE * 1e9 / (t["ms_conv"] * 1e-3) / 1e12,
            "heads_rel_err_vs_fp32_oracle_2_tiles": herr, "heads_within_2e-2": bool(max(herr) <= 2e-2)}


def aux_cross_seam_sparse(local, args, peaks, img, w):
    """The optional cross-seam stage (north_star; not in the reference) in the regime it is made for: detections of bounded
    size spread over the image (100 k boxes of 20-200 px on the 20000^2 grid of 512^2 tiles, edge 64 -> ~49 k seam candidates).
    The bench's own random-weight boxes are heavy-tailed (a fifth of them larger than 1024 px), which sends the stage down its
    general (serial-chunk) route - that number is in aux_cross_seam of the multi-GPU lines."""
    import torch
    from oracle import tiling_np as tl, nms_c
    from yolo3_b200 import post_engine
    pe = post_engine(local)
    side = 20000
    rng = np.random.default_rng(5)

    def boxes(n, ext):
        cx, cy = rng.uniform(0, ext, n), rng.uniform(0, ext, n)
        bw, bh = rng.uniform(20, 200, n), rng.uniform(20, 200, n)
        return np.stack([np.clip(np.round(cx - bw / 2), 0, ext - 1), np.clip(np.round(cy - bh / 2), 0, ext - 1),
                         np.clip(np.round(cx + bw / 2), 0, ext - 1), np.clip(np.round(cy + bh / 2), 0, ext - 1),
                         rng.permutation(n).astype(np.float64) / n * 0.9 + 0.1, np.zeros(n)], 1)
    rows = torch.from_numpy(boxes(100_000, side)).to(torch.device("cuda", local))
    pe.cross_seam_nms(rows, (side, side), TILE, 64, 0.3, number_classes=1)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        out = pe.cross_seam_nms(rows, (side, side), TILE, 64, 0.3, number_classes=1)
        torch.cuda.synchronize()
        ts.append(1e3 * (time.perf_counter() - t0))
    small = boxes(20_000, 8000)
    want = tl.cross_seam_nms(small, (8000, 8000), TILE, 64, 0.3)
    got = pe.cross_seam_nms(small, (8000, 8000), TILE, 64, 0.3, number_classes=1)
    return {"rows_in": 100_000, "rows_out": int(out.shape[0]), "ms_stage_min": float(np.min(ts)), "ms_stage_median": float(np.median(ts)),
            "equals_oracle_stage_on_20k_sample": bool(np.array_equal(got, want))}


def aux_k5(local, rank, world, args, peaks, barrier, dist, dev):
    """BASELINE configs[4]: 608x608x3, NC=80, batch 256 = 32 images per GPU, plain data parallel (no exchange): every rank
    runs forward + decode + filter + NMS on its 32 images; images/s = all ranks' images / max-over-ranks time."""
    import torch
    from yolo3_b200 import Engine, weights as _wts
    B = 32
    e5 = Engine((608, 608, 3), 80, ANCHORS, max_batch=B, device=local)
    w5 = _wts.random_init(3, 80, 3, seed=0, randomize_bn=True)
    e5.load_weights(w5)
    x5 = np.random.default_rng(100 + rank).standard_normal((B, 3, 608, 608)).astype(np.float32)
    calibrate_heads(e5, w5, x5[:4], pass_frac=0.002, nc=80, interior=False)
    xd = torch.from_numpy(x5).to(dev)
    for _ in range(2):
        e5.detect(xd, MIN_BOX, IOU_THR, SCORE_THR)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_rep = 5
    e0.record()
    conv = 0.0
    for _ in range(n_rep):
        e5.detect(xd, MIN_BOX, IOU_THR, SCORE_THR)
        conv += e5.timings()["ms_conv"] / n_rep
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / n_rep], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    out = {"images_per_rank": B, "ranks": world, "ms_per_batch_max_over_ranks": ms, "images_per_s_all_ranks": world * B / (ms * 1e-3),
           "conv_ms_rank0": conv, "conv_tflops_rank0": B * 141.26e9 / (conv * 1e-3) / 1e12,
           "input": "device-resident NCHW fp32 (the reference feeds a tf tensor)", "timed": "forward + decode + filter + NMS + D2H of the boxes"}
    if rank == 0:
        herr = _head_err(e5, w5, (608, 608, 3), 80, x5[:1])
        out.update(heads_rel_err_vs_fp32_oracle_image0=herr, **{"heads_within_2e-2": bool(max(herr) <= 2e-2)})
    del e5
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--image-side", type=int, default=20000)
    ap.add_argument("--edge", type=int, default=64, help="EDGE_EFFECT_RANGE; 64 = BASELINE wording, 96 = reference constant")
    ap.add_argument("--batch", type=int, default=0, help="largest tile batch of the engine (0 = 256, the facade's BATCH_SIZE)")
    ap.add_argument("--cpu-sample-tiles", type=int, default=96)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    side, edge = args.image_side, args.edge
    workload = "inference_tiled %dx%d uint16, %dx%d tiles, %d-px overlap (BASELINE configs[3])" % (side, side, TILE[0], TILE[1], edge)
    base = {"metric": "image megapixels/s, tiled YOLOv3 inference", "unit": "Mpix/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "data": "synthetic", "dtype": "bf16"}
    cfg = {"workload": workload, "tile": list(TILE), "edge_range": edge, "classes": NC, "anchors": len(ANCHORS),
           "weights": "random-init (Keras defaults, randomised BN, calibrated sparse heads)",
           "l2": "inputs larger than L2 (800 MB image, >1 GB of activations per tile batch)"}

    # ------------------------------------------------------------------ reference arm (CPU oracle port)
    if args.impl == "reference":
        if rank != 0:
            return 0
        threads = os.cpu_count() or 1
        img = synthetic_image(min(side, 4096))          # the sample only touches the first tiles
        sample = "first %d tiles per step (torch-CPU fp32 forward + NumPy/C post-processing, oracle port; " \
                 "TensorFlow is not installable here)" % args.cpu_sample_tiles
        for _ in range(min(args.warmup, 1)):
            cpu_baseline_run(img, edge, 1, threads)
        vals, t_all = [], 0.0
        for _ in range(args.steps):
            v, dt = cpu_baseline_run(img, edge, args.cpu_sample_tiles, threads)
            vals.append(v)
            t_all += dt
        v = float(np.mean(vals))
        line = dict(base, impl="reference", value=v, ms_per_step=1e3 * t_all / args.steps, dtype="f32", config=cfg,
                    cpu_baseline={"value": v, "unit": "Mpix/s", "cores": threads, "kind": "port", "sample": sample},
                    e2e={"value": v, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                    gpu_launches=0)
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ B200 arm
    import contextlib
    import hashlib
    import io
    import torch
    import torch.distributed as dist
    import inference_tiled as facade                       # the drop-in module: e2e goes through ITS entry point
    from yolo3_b200 import Engine, infer_tiled_distributed, pinned_copy, shard_range, tile_count

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def sha(a):
        return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]

    img = synthetic_image(side)
    img_host = pinned_copy(img)                            # what imagereader.imread returns for a large image: page-locked
    n_tiles = tile_count(side, side, TILE, edge)
    first, count = shard_range(n_tiles, rank, world)
    if args.batch <= 0:
        # the facade's BATCH_SIZE (inference_tiled.py); the library splits a call's tiles evenly into at least three
        # batches of at most this many tiles.  Measured on one B200 (power-capped step): 64 -> 1164, 128 -> 1224,
        # 192 -> 1244, 256 -> 1252, 384 -> 1247 Mpix/s
        args.batch = min(256, max(1, count))
    cfg.update(tiles=n_tiles, tiles_this_rank=count, tile_batch=args.batch, parallelism="tile-sharded x%d" % world)

    eng = Engine(TILE + (1,), NC, ANCHORS, max_batch=args.batch, device=local)
    w = bench_weights()
    eng.load_weights(w)
    calibrate_heads(eng, w, eng.tiles_normalized(np.ascontiguousarray(img[:1024, :2048]), TILE, edge, 0, min(4, args.batch)))
    img_dev = torch.from_numpy(img.view(np.int16)).to(dev).view(torch.uint16)      # the image resident in HBM
    model_obj = type("LoadedModelLike", (), {"engine": eng})()                      # what model.load_saved_model returns carries .engine
    facade.EDGE_EFFECT_RANGE = edge
    if world > 1:
        eng.comm_init()

    def step(resident):
        if resident:      # image already in HBM: the library call directly (the facade takes host arrays, as the reference does)
            if world > 1:
                return infer_tiled_distributed(eng, img_dev, TILE, MIN_BOX, edge, IOU_THR, SCORE_THR)
            return eng.infer_tiled(img_dev, TILE, MIN_BOX, edge, IOU_THR, SCORE_THR, out_device=dev)
        # end to end: the drop-in entry point, host image in, host float64 [n,6] out
        with contextlib.redirect_stdout(io.StringIO()):
            return facade.inference_image_tiled(model_obj, img_host, list(TILE), MIN_BOX)

    def timed(resident):
        for _ in range(args.warmup):
            out = step(resident)
        barrier()
        k0 = eng.timings()["kernels_launched"]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        conv_ms = 0.0
        tiles_done = 0
        stage = {}
        for _ in range(args.steps):
            out = step(resident)
            t = eng.timings()
            conv_ms += t["ms_conv"]
            tiles_done += t["tiles"]
            for k in ("ms_h2d", "ms_prep", "ms_conv", "ms_decode", "ms_nms", "ms_stitch", "ms_d2h", "ms_comm"):
                stage[k] = stage.get(k, 0.0) + t[k] / args.steps
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        t = eng.timings()
        t["tiles_per_step"] = tiles_done / args.steps            # this rank's shard (follows its measured throughput when sharded)
        return float(ms.item()) / args.steps, out, conv_ms / args.steps, t["kernels_launched"] - k0, stage, t

    sampler = ClockSampler(local)
    sampler.start()
    ms_res, out_res, conv_ms, launches, stages, tlast = timed(True)
    ms_e2e, out_e2e, _, _, stages_e2e, tlast_e2e = timed(False)
    clocks = sampler.summary()

    out_res_np = out_res.cpu().numpy() if hasattr(out_res, "cpu") else np.asarray(out_res)
    n_boxes = int(out_res_np.shape[0])
    mpix = side * side / 1e6
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "measured sustained (MEASURED_PEAKS.json)" if peaks else "fallback"
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    count = tlast["tiles_per_step"]                            # tiles this rank really ran per step
    cfg["tiles_this_rank_measured"] = count
    conv_tf = count * CONV_GF_PER_TILE * 1e9 / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    h2d = int(side * side * 2 / world) if world > 1 else side * side * 2
    # traffic of the dominant kernel: dram bytes of ONE launch from an `ncu --set full` capture, parsed into profiles/ by
    # tools/ncu_summary.py (not measurable inside a timed run); null when no capture has been committed
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")))
    except Exception:
        pass
    # share of the conv time by kernel class, from the library's per-layer timer at this batch size (isolated layers)
    classes = {}
    try:
        rows = [r.split(",") for r in eng.profile_layers(min(args.batch, 128), 2).strip().splitlines()[1:]]
        tot = sum(float(r[13]) for r in rows)
        for r in rows:
            k, cout, cin = int(r[2]), int(r[5]), int(r[4])
            cls = ("k_conv_tc2h<256> (3x3, Cout % 256 == 0)" if k == 3 and cout % 256 == 0 and cin >= 128 else
                   "k_conv_halo / k_stem_conv1 (3x3, Cin <= 64)" if k == 3 and cin <= 64 else
                   "k_conv_tc2 (1x1 Cout >= 128, 3x3 Cout 128)" if cout >= 128 else "k_conv_tc (Cout < 128, heads)")
            classes[cls] = classes.get(cls, 0.0) + float(r[13]) / tot
    except Exception:
        pass
    line = dict(base, value=mpix / (ms_res * 1e-3), ms_per_step=ms_res, config=cfg, clocks=clocks,
                e2e={"value": mpix / (ms_e2e * 1e-3), "unit": "Mpix/s", "h2d_bytes_per_step": h2d,
                     "d2h_bytes_per_step": n_boxes * 48, "ms_per_step": ms_e2e,
                     "api": "inference_tiled.inference_image_tiled(model, pinned host image, [512,512], 32)"},
                gpu_launches=int(launches),
                roofline={"bound": "tensor", "kernel": "conv stack of the step on this rank (75 Conv2D per tile batch); dominant class by time: "
                          + (max(classes, key=classes.get) if classes else "k_conv_tc2h<256>"),
                          "achieved": conv_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": conv_tf / peak_tf,
                          "peak_source": peak_src, "frac_of_burst": conv_tf / float(peaks.get("bf16_tflops", 1664.7)),
                          "traffic": traffic.get("dram_bytes_per_launch"), "traffic_launch": traffic.get("launch"),
                          "traffic_algorithmic_bytes": traffic.get("algorithmic_bytes_per_launch"), "traffic_source": traffic.get("source"),
                          "kernel_class_time_share_isolated": classes,
                          "flops_per_step_this_rank": count * CONV_GF_PER_TILE * 1e9, "conv_ms_per_step": conv_ms},
                stages_ms=stages, stages_ms_e2e=stages_e2e, lib_ms_total_e2e=tlast_e2e["ms_total"], boxes=n_boxes,
                candidates_per_step=int(tlast["candidates"]),
                network_input_mpix_per_s=n_tiles * TILE[0] * TILE[1] / 1e6 / (ms_res * 1e-3))
    line["e2e_output_equals_resident_output"] = bool(np.array_equal(np.asarray(out_e2e), out_res_np))

    # ---- K5 (BASELINE configs[4]): 608x608x3, NC=80, 32 images per GPU, data parallel - every rank, aggregate images/s
    try:
        line_k5 = aux_k5(local, rank, world, args, peaks, barrier, dist if world > 1 else None, dev)
    except Exception as ex:
        line_k5 = {"error": str(ex)[:200]}
    if world > 1:
        # the N-rank gathered output must be the 1-rank output, row for row (rank 0 runs the whole image alone once)
        if rank == 0:
            single = eng.infer_tiled(img_dev, TILE, MIN_BOX, edge, IOU_THR, SCORE_THR)
            line["rank_output_matches_single"] = bool(np.array_equal(single, out_res_np))
            line["output_sha"] = {"sharded": sha(out_res_np), "single_gpu": sha(single)}
            # cross-seam stage (north_star; not in the reference): timed once on the sharded path
            t0 = time.perf_counter()
        seam = infer_tiled_distributed(eng, img_dev, TILE, MIN_BOX, edge, IOU_THR, SCORE_THR, cross_seam=True)
        if rank == 0:
            torch.cuda.synchronize()
            line["aux_cross_seam"] = {"ms_step_with_stage": 1e3 * (time.perf_counter() - t0), "rows_in": n_boxes, "rows_out": int(seam.shape[0]),
                                      "ms_comm_incl_stage": eng.timings()["ms_comm"]}
    if rank == 0:
        line["aux_k5_608_b32_per_gpu_nc80"] = line_k5
        if world == 1:
            del eng
            torch.cuda.empty_cache()
            for name, fn in (("aux_nms_k3", aux_k3), ("aux_k2_416_b64_nc80", aux_k2), ("aux_k1_416_b1_nc80", aux_k1), ("aux_k4_edge96", aux_k4_edge96),
                             ("aux_cross_seam_sparse", aux_cross_seam_sparse)):
                try:
                    line[name] = fn(local, args, peaks, img, w)
                except Exception as ex:                      # auxiliary only - never fail the headline line
                    line[name] = {"error": "%s: %s" % (type(ex).__name__, str(ex)[:200])}
                torch.cuda.empty_cache()
            threads = os.cpu_count() or 1
            v, dt = cpu_baseline_run(img[:4096, :4096], edge, args.cpu_sample_tiles, threads)
            line["cpu_baseline"] = {"value": v, "unit": "Mpix/s", "cores": threads, "kind": "port",
                                    "sample": "first %d tiles of the same workload, %.1f s (oracle port: torch-CPU fp32 "
                                              "forward + NumPy/C post-processing; TensorFlow not installable)" % (args.cpu_sample_tiles, dt)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
