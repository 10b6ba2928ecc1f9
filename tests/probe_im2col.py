"""Hardware probe (GPU): semantics of im2col-mode TMA loads vs a NumPy im2col."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "object-detection-yolov3_b200"))
from yolo3_b200 import _lib  # noqa: E402
from yolo3_b200 import post_engine
from yolo3_b200._lib import check
eng = post_engine(0)


def bf16_bits(a):
    return (a.astype(np.float32).view(np.uint32) >> 16).astype(np.uint16)


def from_bits(b):
    return (b.astype(np.uint32) << 16).view(np.float32)


def unswizzle(raw):                       # raw [128][64] bf16 bits in 128B-swizzled rows
    t = raw.reshape(128, 8, 8)
    out = np.empty_like(t)
    for r in range(128):
        for j in range(8):
            out[r, j] = t[r, j ^ (r & 7)]
    return out.reshape(128, 64)


def expected(x, stride, pad_lo, k, m0, tap_w, tap_h, c0):
    N, H, W, C = x.shape
    Ho, Wo = H // stride, W // stride
    out = np.zeros((128, 64), np.float32)
    for r in range(128):
        m = m0 + r
        n, rem = divmod(m, Ho * Wo)
        ho, wo = divmod(rem, Wo)
        if n >= N:
            continue
        hi, wi = ho * stride - pad_lo + tap_h, wo * stride - pad_lo + tap_w
        if 0 <= hi < H and 0 <= wi < W:
            out[r] = x[n, hi, wi, c0:c0 + 64]
    return out


rng = np.random.default_rng(0)
for (N, H, W, C, stride, pad_lo, pad_hi) in [(3, 13, 13, 128, 1, 1, 1), (2, 12, 12, 64, 2, 0, 1), (2, 26, 26, 64, 1, 1, 1)]:
    x = rng.integers(-100, 100, (N, H, W, C)).astype(np.float32)
    Ho, Wo = H // stride, W // stride
    cases = []
    for m0 in (0, 5, 128, 2 * Ho * Wo - 40 if N > 2 else Ho * Wo - 40, N * Ho * Wo - 60):
        for (tw, th) in ((0, 0), (1, 1), (2, 2), (2, 0)):
            n, rem = divmod(m0, Ho * Wo)
            ho, wo = divmod(rem, Wo)
            cases.append((0 if C == 64 else 64, wo * stride - pad_lo, ho * stride - pad_lo, n, tw, th, m0))
    probes = np.array([c[:6] for c in cases], np.int32)
    out = np.zeros((len(cases), 128, 64), np.uint16)
    bits = bf16_bits(x)
    check(_lib.load_probe().y3_debug_im2col(eng.h, bits.ctypes.data, N, H, W, C, stride, pad_lo, pad_hi, 3, probes.ctypes.data, len(cases),
                                  out.ctypes.data), eng.h)
    bad = 0
    for i, c in enumerate(cases):
        got = from_bits(unswizzle(out[i]))
        want = expected(x, stride, pad_lo, 3, c[6], c[4], c[5], c[0])
        rows_ok = int((got == want).all(axis=1).sum())
        nan_rows = int(np.isnan(got).any(axis=1).sum())
        if rows_ok != 128:
            bad += 1
            first = int(np.argmin((got == want).all(axis=1)))
            print("  MISMATCH m0=%d tap=(%d,%d): %d/128 rows exact, %d rows untouched(NaN), first bad row %d" % (c[6], c[4], c[5], rows_ok, nan_rows, first))
    print("N%d H%d W%d C%d stride %d pad (%d,%d): %d/%d probes exact" % (N, H, W, C, stride, pad_lo, pad_hi, len(cases) - bad, len(cases)))
