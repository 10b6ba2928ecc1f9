"""Seeded synthetic inputs shared by gen_golden.py, the tests and bench.py (ORACLE side).

SURVEY.md section 8(d) defines the concrete inputs; this file is their single source.
"""
import numpy as np

from . import postproc_np as pp

F32 = np.float32


def boxes_on_canvas(n, canvas, wh_lo, wh_hi, rng):
    c = rng.uniform(0, canvas, (n, 2))
    wh = rng.uniform(wh_lo, wh_hi, (n, 2))
    return np.concatenate([c - wh / 2, c + wh / 2], 1).astype(F32)


def nms_case(n, canvas, seed, wh=(33, 300)):
    """K3-style single class: n boxes on a canvas^2, tie-free fp32 scores."""
    rng = np.random.default_rng(seed)
    b = boxes_on_canvas(n, canvas, wh[0], wh[1], rng)
    s = pp.make_tie_free_scores(n, rng)
    return b, s


def k3_single_class(n=200_000, seed=0):
    return nms_case(n, 2000, seed)


def multiclass_case(n, nc, canvas, seed, dominant_only=True):
    """K3 80-class style: obj = 1 and cls = score^2 exactly representable-ish, one dominant class
    per box >= 0.1 and every other class < 0.1 (dominant_only) or every class passing (heavy).
    Scores sqrt(cls*obj) are made tie-free per class by construction (distinct cls values)."""
    rng = np.random.default_rng(seed)
    b = boxes_on_canvas(n, canvas, 33, 300, rng)
    obj = np.ones((n, 1), F32)
    if dominant_only:
        cls = (rng.uniform(0.0, 0.009, (n, nc))).astype(F32)          # sqrt < 0.095
        dom = rng.integers(0, nc, n)
        lad = pp.make_tie_free_scores(n, rng, lo=0.02, hi=0.98)        # distinct -> distinct sqrt? checked below
        cls[np.arange(n), dom] = lad
    else:
        lad = pp.make_tie_free_scores(n * nc, rng, lo=0.02, hi=0.98)
        cls = lad.reshape(n, nc).copy()
    sc = pp.blended_scores(obj, cls)
    for c in range(nc):
        v = sc[:, c][sc[:, c] >= F32(0.1)]
        if np.unique(v).size != v.size:                                # nudge the rare collision
            raise AssertionError("score collision in class %d - change the seed" % c)
    return b, obj, cls


def degenerate_case(seed=11, n=500):
    """zero-area, duplicate, inf and NaN boxes (0/0 IoU -> NaN -> suppressed)."""
    rng = np.random.default_rng(seed)
    b = rng.uniform(0, 50, (n, 4)).astype(F32)
    b[:, 2:] = b[:, :2] + rng.uniform(0, 30, (n, 2)).astype(F32)
    b[::7, 2:] = b[::7, :2]
    b[5] = b[12]
    b[40] = [np.inf, 0, np.inf, 5]
    b[41] = [np.nan, 1, 3, 4]
    s = pp.make_tie_free_scores(n, rng)
    return b, s


def synthetic_image(h, w, c, dtype, seed, blobs=0):
    rng = np.random.default_rng(seed)
    info = np.iinfo(dtype) if np.issubdtype(dtype, np.integer) else None
    if info is not None:
        img = rng.integers(0, min(info.max, 65535), (h, w, c), dtype=dtype)
    else:
        img = rng.standard_normal((h, w, c)).astype(dtype)
    for _ in range(blobs):
        y, x = rng.integers(0, h), rng.integers(0, w)
        r = int(rng.integers(8, 40))
        img[max(0, y - r):y + r, max(0, x - r):x + r] = info.max if info is not None else 5.0
    return img


class FakeDetector:
    """Deterministic stand-in for yolo_model(batch, training=False) used to pin the tile
    pipeline (a12, a17) independently of the network: the detections of a tile are a pure
    function of the tile's normalised pixels, so a wrong slice / reflect / normalise shows up."""

    def __init__(self, n_boxes, nc, tile_hw, seed=5):
        rng = np.random.default_rng(seed)
        th, tw = tile_hw
        self.n, self.nc, self.th, self.tw = n_boxes, nc, th, tw
        self.cy = rng.integers(0, th, n_boxes)
        self.cx = rng.integers(0, tw, n_boxes)
        self.wh = rng.uniform(20, 90, (n_boxes, 2)).astype(F32)
        self.obj = rng.uniform(0.0, 1.0, n_boxes).astype(F32)
        self.cls = rng.uniform(0.0, 1.0, (n_boxes, nc)).astype(F32)

    def __call__(self, batch, training=False):
        x = np.asarray(batch, dtype=F32)[0]                       # [C,H,W]
        v = x[0, self.cy, self.cx]                                # pixel under each box centre
        jit = np.tanh(v).astype(F32)                              # in (-1,1), depends on slice+normalise
        cx = self.cx.astype(F32) + F32(3.0) * jit
        cy = self.cy.astype(F32) - F32(2.0) * jit
        w = self.wh[:, 0] * (F32(1.0) + F32(0.25) * jit)
        h = self.wh[:, 1] * (F32(1.0) - F32(0.25) * jit)
        obj = np.clip(self.obj + F32(0.1) * jit, 0, 1).astype(F32)
        out = np.concatenate([(cx - w / 2)[:, None], (cy - h / 2)[:, None], (cx + w / 2)[:, None],
                              (cy + h / 2)[:, None], obj[:, None], self.cls], 1).astype(F32)
        return out[None]


# (h, w, c, dtype, tile, edge) - tile slicing fixtures of tests/golden/tiling.npz
TILE_CASES = dict(u16=(700, 900, 1, np.uint16, (512, 512), 96), u8rgb=(520, 1100, 3, np.uint8, (256, 320), 64),
                  small=(300, 280, 1, np.uint16, (512, 512), 96), wide=(400, 1500, 1, np.uint16, (512, 512), 96))
# (h, w, c, dtype, tile, edge, n_boxes, nc, min_box) - tests/golden/tiled_pipeline.npz
PIPE_CASES = dict(e96=(1200, 1500, 1, np.uint16, (512, 512), 96, 900, 2, 32),
                  e64=(1000, 1300, 1, np.uint16, (512, 512), 64, 700, 1, 32),
                  rgb=(700, 640, 3, np.uint8, (256, 256), 32, 400, 3, 24),
                  one=(300, 280, 1, np.uint16, (512, 512), 96, 300, 1, 32))


def merge_case(n, canvas, seed, wh=(10, 60)):
    """integer-coordinate boxes + distinct scores for bbox_utils.union_all_overlapping_bb (exact in fp32 and fp64)"""
    rng = np.random.default_rng(seed)
    c = rng.integers(0, canvas, (n, 2))
    half = rng.integers(wh[0] // 2, wh[1] // 2, (n, 2))
    boxes = np.concatenate([c - half, c + half], axis=1).astype(np.int64)
    scores = (rng.permutation(n).astype(np.float64) + 1.0) / (n + 1.0)
    return boxes, scores
