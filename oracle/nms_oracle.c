/* ORACLE - test infrastructure only (see oracle/__init__.py).  Plain-C restatement of the
 * reference's greedy NMS so that full-size (200k box) bit-exact parity finishes in seconds.
 *
 * Follows /root/reference/bbox_utils.py:
 *   compute_iou        200-214   inter = max(yb-yt,0)*max(xr-xl,0); union=(a+b)-inter; iou=inter/union
 *   single_class_nms   217-237   order by score desc; pop best, keep, survivors = iou <= thr
 *   per_class_nms      240-271   score = sqrt(cls*obj) >= thr per class, class-major output
 * fp32 throughout, IEEE division, NaN compares false (=> suppressed), no FMA contraction
 * (built with -ffp-contract=off).  np.maximum/np.minimum propagate NaN, hence npmax/npmin.
 * Tie rule: score desc, index asc (the reference's argsort()[::-1] is unpinned on ties).
 * Pinned against the reference run verbatim by tests/test_oracle_pinned.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline float npmax(float a, float b) { return (a >= b || a != a) ? a : b; }
static inline float npmin(float a, float b) { return (a <= b || a != a) ? a : b; }

typedef struct { float s; int32_t i; } sk_t;

static int cmp_desc(const void* pa, const void* pb) {
    const sk_t* a = (const sk_t*)pa; const sk_t* b = (const sk_t*)pb;
    if (a->s > b->s) return -1;
    if (a->s < b->s) return 1;
    return (a->i > b->i) - (a->i < b->i);
}

/* boxes [m,4] (x0,y0,x1,y1), scores [m]; keep receives <= m indices; returns count.
 * The alive set is kept as compacted SoA arrays (like the reference's fancy-indexed copies)
 * so that the IoU pass streams and vectorises; lanes stay IEEE-exact. */
int64_t y3o_single_class_nms(const float* boxes, const float* scores, int64_t m, float thr, int32_t* keep) {
    if (m <= 0) return 0;
    sk_t* ord = (sk_t*)malloc(sizeof(sk_t) * (size_t)m);
    float* buf = (float*)malloc(sizeof(float) * 5 * (size_t)m);
    float* X0 = buf, *Y0 = buf + m, *X1 = buf + 2*m, *Y1 = buf + 3*m, *AR = buf + 4*m;
    int32_t* ID = (int32_t*)malloc(sizeof(int32_t) * (size_t)m);
    unsigned char* ok = (unsigned char*)malloc((size_t)m);
    for (int64_t i = 0; i < m; ++i) { ord[i].s = scores[i]; ord[i].i = (int32_t)i; }
    qsort(ord, (size_t)m, sizeof(sk_t), cmp_desc);
    for (int64_t r = 0; r < m; ++r) {
        const int32_t i = ord[r].i;
        X0[r] = boxes[4*i]; Y0[r] = boxes[4*i+1]; X1[r] = boxes[4*i+2]; Y1[r] = boxes[4*i+3];
        const float w = X1[r] - X0[r], h = Y1[r] - Y0[r];
        AR[r] = w * h;
        ID[r] = i;
    }
    int64_t lo = 0, hi = m, n_keep = 0;       /* alive = [lo, hi) */
    while (lo < hi) {
        keep[n_keep++] = ID[lo];
        const float bx0 = X0[lo], by0 = Y0[lo], bx1 = X1[lo], by1 = Y1[lo], ba = AR[lo];
        ++lo;
        for (int64_t r = lo; r < hi; ++r) {
            const float xl = npmax(bx0, X0[r]);
            const float yt = npmax(by0, Y0[r]);
            const float xr = npmin(bx1, X1[r]);
            const float yb = npmin(by1, Y1[r]);
            const float dh = npmax(yb - yt, 0.0f);
            const float dw = npmax(xr - xl, 0.0f);
            const float inter = dh * dw;
            const float sum = ba + AR[r];
            const float uni = sum - inter;
            const float iou = inter / uni;
            ok[r] = (unsigned char)(iou <= thr);
        }
        int64_t w = lo;
        for (int64_t r = lo; r < hi; ++r) {
            if (ok[r]) {
                X0[w] = X0[r]; Y0[w] = Y0[r]; X1[w] = X1[r]; Y1[w] = Y1[r]; AR[w] = AR[r]; ID[w] = ID[r];
                ++w;
            }
        }
        hi = w;
    }
    free(ord); free(buf); free(ID); free(ok);
    return n_keep;
}

/* boxes [n,4], obj [n], cls [n,nc] -> out_boxes [cap,4], out_scores [cap], out_labels [cap]; returns k
 * (or -needed when cap is too small).  Class-major, score-descending inside a class. */
int64_t y3o_per_class_nms(const float* boxes, const float* obj, const float* cls, int64_t n, int32_t nc,
                          float iou_thr, float score_thr, float* out_boxes, float* out_scores,
                          int32_t* out_labels, int64_t cap) {
    float* fb = (float*)malloc(sizeof(float) * 4 * (size_t)(n > 0 ? n : 1));
    float* fs = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
    int32_t* kp = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    int64_t k = 0;
    for (int32_t c = 0; c < nc; ++c) {
        int64_t m = 0;
        for (int64_t i = 0; i < n; ++i) {
            const float p = cls[i * nc + c] * obj[i];
            const float s = sqrtf(p);
            if (s >= score_thr) { memcpy(fb + 4*m, boxes + 4*i, 16); fs[m] = s; ++m; }
        }
        if (!m) continue;
        const int64_t nk = y3o_single_class_nms(fb, fs, m, iou_thr, kp);
        for (int64_t t = 0; t < nk; ++t, ++k) {
            if (k < cap) {
                memcpy(out_boxes + 4*k, fb + 4*kp[t], 16);
                out_scores[k] = fs[kp[t]];
                out_labels[k] = c;
            }
        }
    }
    free(fb); free(fs); free(kp);
    return k <= cap ? k : -k;
}
