// nms.cu - confidence thresholding, warp-ballot stream compaction, (class,score) ordering and
// exact greedy per-class NMS on the device.
//
// Reference behaviour reproduced bit for bit (fp32, NumPy operand order, no FMA contraction):
//   filter_small_boxes  bbox_utils.py:274-281   (w > min) & (h > min), strict
//   per_class_nms       bbox_utils.py:240-271   score = sqrt(cls*obj) >= float32(thr), class-major output
//   single_class_nms    bbox_utils.py:217-237   greedy, survivors are (iou <= thr); NaN => suppressed
//   compute_iou         bbox_utils.py:200-214
// Ordering rule: score descending, then input row ascending (the reference's argsort()[::-1] is
// unpinned on ties - SURVEY.md Q11 - this is the one documented deviation).
//
// This file holds (1) the threshold + compaction kernels shared by both pipelines and (2) the global-sort pipeline.
// The default pipeline is the segmented, synchronisation-free one in nms_seg.cu; PostProc::enqueue() routes:
//   every (image, class) segment can fit one CTA's shared memory (<= 24576 boxes)  -> nms_seg.cu
//   otherwise the largest segment is read back once: still <= 24576                 -> nms_seg.cu
//                                                    larger (e.g. 200 k boxes, 1 class) -> global sort (below)
//
// Threshold + compaction (all routes):
//   k_candidates          one thread per (row, class) of decoded rows (also raw heads: the pre-round-2 form, A/B)
//   k_candidates_heads1   raw heads, one class: one lane per row
//   k_live_rows + k_candidates_rows   raw heads, many classes: list the rows whose objectness can pass, then one warp per
//                         listed row.  Candidates = 64-bit key [segment | ~orderable(score) | row], the global row, the
//                         decoded box and (segmented route) the slot inside the segment
// Global-sort pipeline (all on ctx->stream, host-synchronised):
//   k_rs_*            own stable LSD radix sort (8 bits / pass) on the used key bits only
//   k_gather_sorted   boxes / areas of the sorted candidates as SoA (coalesced for the sweeps)
//   k_seg_offsets     segment (= image x class) boundaries by binary search on the sorted keys
//   k_nms_small / k_nms_segments    one warp / one CTA per segment: per 512-box chunk a shared-memory IoU bitmask
//                     (512 x 8 u64), a serial suppression sweep over it, then the chunk's kept
//                     boxes are applied to the rest of the segment
//   k_nms_resolve_next / k_nms_apply_from   the same two phases as separate launches for very large
//                     segments, so that the apply phase uses the whole GPU
//   k_count / k_scan / k_scatter  ordered compaction of the keep flags into the output arrays
// The n x n/64 bitmask is never materialised in HBM.
#include "nms_common.cuh"

#include <math.h>
#include <stdlib.h>
#include <algorithm>

namespace y3 {

// Where the candidates go.  seg_cnt / slot (optional) feed the segmented pipeline (nms_seg.cu): every candidate also
// takes the next slot of its (image, class) segment, so that one scan + one scatter bin the list by segment.
struct CandOut {
    uint64_t* keys;
    uint32_t* vals;
    unsigned long long* counter;
    int64_t cap;
    int* seg_cnt;
    uint32_t* slot;
    float4* box;          // optional: the candidate's decoded box (already computed for the small-box filter), so that the
                          // NMS kernels of the segmented pipeline load 16 bytes instead of decoding the row again
};

// All 32 lanes call this (converged); lanes with `pass` append (key, val).  One atomic per warp for the list position
// and one per distinct segment among the passing lanes (match-any aggregation) for the slot.
__device__ __forceinline__ void emit_candidates(const CandOut& O, bool pass, uint64_t key, uint32_t val, uint32_t seg, int lane,
                                                const float4 bx = make_float4(0.f, 0.f, 0.f, 0.f)) {
    const unsigned m = __ballot_sync(0xffffffffu, pass);
    if (!m) return;
    unsigned long long base = 0;
    const int leader = __ffs(m) - 1;
    if (lane == leader) base = atomicAdd(O.counter, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (pass) {
        const int64_t pos = (int64_t)base + __popc(m & ((1u << lane) - 1u));
        uint32_t sl = 0;
        if (O.seg_cnt) {
            const unsigned peers = __match_any_sync(m, seg);
            const int l2 = __ffs(peers) - 1;
            int b2 = 0;
            if (lane == l2) b2 = atomicAdd(O.seg_cnt + seg, __popc(peers));
            b2 = __shfl_sync(peers, b2, l2);
            sl = (uint32_t)(b2 + __popc(peers & ((1u << lane) - 1u)));
        }
        if (pos < O.cap) {
            O.keys[pos] = key;
            O.vals[pos] = val;
            if (O.slot) O.slot[pos] = sl;
            if (O.box) O.box[pos] = bx;
        }
    }
}

__device__ __forceinline__ uint64_t make_key(const KeyLayout& kl, uint64_t seg, float s, uint32_t row) {
    return (seg << kl.seg_shift) | ((uint64_t)(~orderable(s) - kl.score_base) << kl.row_bits) | (uint64_t)row;
}

// One thread per (row, class) of decoded rows (or of raw heads: the pre-round-2 form, kept for A/B).
// IDX = uint32_t when the element count fits (the common case): 64-bit divisions are ~10x slower
template <typename IDX>
__global__ void __launch_bounds__(256)
k_candidates(CandSource src, KeyLayout kl, int64_t total, float logit_floor, CandOut O) {
    const IDX stride = (IDX)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    // total is rounded up to a multiple of 32 by the loop bound so that ballots stay converged
    const IDX total_r = (IDX)((total + 31) & ~(int64_t)31);
    const IDX nc = (IDX)src.nc, rpi = (IDX)src.rows_per_image;
    const float thr = src.score_thr;
    for (IDX e = (IDX)blockIdx.x * blockDim.x + threadIdx.x; e < total_r; e += stride) {
        bool pass = false;
        uint64_t key = 0;
        uint32_t val = 0, seg = 0;
        float4 cbox = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < (IDX)total) {
            const IDX grow = e / nc;                         // global row = img * rows + row
            const int c = (int)(e - grow * nc);
            const IDX img = grow / rpi;
            const IDX row = grow - img * rpi;
            float s = 0.f;
            if (src.from_heads) {
                // fused decode: sigmoid of the objectness / class logits straight from the head (model.py:184-185)
                int sc, cell, a;
                const float* ho;
                const float* hp = head_row(src.dec, (int)img, (int)row, &sc, &cell, &a, &ho);
                const float lo = __ldg(ho), lc = __ldg(hp + 4 + c);
                // exact conservative pre-filter: score^2 = obj*cls <= min(obj, cls); a logit below
                // logit(thr^2) - margin cannot reach the threshold, so both sigmoids are skipped
                if (lo >= logit_floor && lc >= logit_floor) {
                    const float o = sigmoid_f(lo);
                    const float p = sigmoid_f(lc);
                    s = __fsqrt_rn(__fmul_rn(p, o));
                    pass = (s >= thr);
                    if (pass && (src.filter_small || O.box)) {
                        cbox = decode_box(src.dec, hp, sc, cell, a);
                        if (src.filter_small) pass = (__fsub_rn(cbox.z, cbox.x) > src.min_size) && (__fsub_rn(cbox.w, cbox.y) > src.min_size);
                    }
                }
            } else {
                const float p = __ldg(src.cls + (int64_t)grow * src.cls_stride + c);
                if (src.raw_scores) {
                    s = p;
                    pass = src.row_mask ? (src.row_mask[grow] != 0) : true;
                } else {
                    const float o = src.obj ? __ldg(src.obj + (int64_t)grow * src.obj_stride) : 1.0f;
                    s = __fsqrt_rn(__fmul_rn(p, o));             // np.sqrt(class_probs * objectness)
                    pass = (s >= thr);                           // NaN -> false
                }
                if (pass && (src.filter_small || O.box)) {
                    const float* b = src.box + (int64_t)grow * src.box_stride;
                    cbox = make_float4(__ldg(b + 0), __ldg(b + 1), __ldg(b + 2), __ldg(b + 3));
                    if (src.filter_small) pass = (__fsub_rn(cbox.z, cbox.x) > src.min_size) && (__fsub_rn(cbox.w, cbox.y) > src.min_size);
                }
            }
            if (pass && src.row_seg) {
                const int sg = src.row_seg[grow];
                pass = sg >= 0 && sg < src.n_seg_override;       // a label outside the class range takes no part
            }
            if (pass) {
                seg = src.row_seg ? (uint32_t)src.row_seg[grow] : (uint32_t)((uint64_t)img * src.nc + (uint64_t)c);
                key = make_key(kl, seg, s, (uint32_t)row);
                val = (uint32_t)grow;
            }
        }
        emit_candidates(O, pass, key, val, seg, lane, cbox);
    }
}

// Fused decode + threshold + compaction straight from the raw fp32 heads (same exact arithmetic as k_candidates).
// One lane per row: 32 independent objectness loads in flight per warp; a row whose objectness logit cannot reach the
// threshold costs this one 4-byte load (score^2 = obj*cls <= obj).
//   one class   : k_candidates_heads1 - the lane finishes its own row (class logit, score, size filter, emit);
//   many classes: k_live_rows appends the rows that can still pass to a list, k_candidates_rows then gives every
//                 listed row to ONE WARP (grid-stride over the device-side list length): the 32 lanes read the row's
//                 class logits side by side, ONE atomic reserves the row's list positions and every passing class takes
//                 its segment slot with an independent atomic.  (Round 1 walked rows in place: the rows that pass
//                 cluster spatially, so a few warps did most of the work serially - 18 % warps active, 0.10 ms on K2.)
__global__ void __launch_bounds__(256)
k_candidates_heads1(CandSource src, KeyLayout kl, float logit_floor, CandOut O) {
    pdl_chain_sync();
    const int lane = threadIdx.x & 31;
    const uint32_t rpi = (uint32_t)src.rows_per_image;
    const uint32_t rows_total = rpi * (uint32_t)src.n_images;
    const uint32_t warp_g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    const float thr = src.score_thr;
    for (uint32_t base = warp_g * 32u; base < rows_total; base += n_warps * 32u) {
        const uint32_t grow = base + (uint32_t)lane;
        bool pass = false;
        float s = 0.f;
        uint32_t img = 0, row = 0;
        float4 cbox = make_float4(0.f, 0.f, 0.f, 0.f);
        if (grow < rows_total) {
            img = grow / rpi;
            row = grow - img * rpi;
            int sc, cell, a;
            const float* ho;
            const float* hp = head_row(src.dec, (int)img, (int)row, &sc, &cell, &a, &ho);
            const float lo = __ldg(ho);
            if (lo >= logit_floor) {
                const float lc = __ldg(hp + 4);
                if (lc >= logit_floor) {
                    s = __fsqrt_rn(__fmul_rn(sigmoid_f(lc), sigmoid_f(lo)));
                    pass = (s >= thr);
                    if (pass && (src.filter_small || O.box)) {
                        cbox = decode_box(src.dec, hp, sc, cell, a);
                        if (src.filter_small) pass = (__fsub_rn(cbox.z, cbox.x) > src.min_size) && (__fsub_rn(cbox.w, cbox.y) > src.min_size);
                    }
                }
            }
        }
        emit_candidates(O, pass, make_key(kl, img, s, row), grow, img, lane, cbox);
    }
}

__global__ void __launch_bounds__(256)
k_live_rows(CandSource src, float logit_floor, uint32_t* __restrict__ live, unsigned int* __restrict__ n_live) {
    pdl_chain_sync();
    const int lane = threadIdx.x & 31;
    const uint32_t rpi = (uint32_t)src.rows_per_image;
    const uint32_t rows_total = rpi * (uint32_t)src.n_images;
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t rows_r = (rows_total + 31u) & ~31u;
    for (uint32_t grow = blockIdx.x * blockDim.x + threadIdx.x; grow < rows_r; grow += stride) {
        bool alive = false;
        if (grow < rows_total) {
            const uint32_t img = grow / rpi;
            int sc, cell, a;
            const float* ho;
            head_row(src.dec, (int)img, (int)(grow - img * rpi), &sc, &cell, &a, &ho);
            alive = __ldg(ho) >= logit_floor;
        }
        const unsigned m = __ballot_sync(0xffffffffu, alive);
        if (m) {
            unsigned base = 0;
            const int leader = __ffs(m) - 1;
            if (lane == leader) base = atomicAdd(n_live, (unsigned)__popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (alive) live[base + __popc(m & ((1u << lane) - 1u))] = grow;
        }
    }
}

static constexpr int ROW_IT = 4;     // class chunks of 32 handled per emission group (128 classes)

__global__ void __launch_bounds__(256)
k_candidates_rows(CandSource src, KeyLayout kl, float logit_floor, const uint32_t* __restrict__ live,
                  const unsigned int* __restrict__ n_live, CandOut O) {
    pdl_chain_sync();
    const int lane = threadIdx.x & 31;
    const uint32_t rpi = (uint32_t)src.rows_per_image;
    const uint32_t warp_g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    const int nc = src.nc;
    const float thr = src.score_thr;
    const uint32_t n = *n_live;
    for (uint32_t li = warp_g; li < n; li += n_warps) {
        const uint32_t grow = live[li];
        const uint32_t img = grow / rpi, row = grow - img * rpi;
        int sc, cell, a;
        const float* ho;
        const float* hp = head_row(src.dec, (int)img, (int)row, &sc, &cell, &a, &ho);
        const float o = sigmoid_f(__ldg(ho));
        bool size_known = false, size_ok = true;
        float4 cbox = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c00 = 0; c00 < nc; c00 += 32 * ROW_IT) {
            float lc[ROW_IT];
#pragma unroll
            for (int it = 0; it < ROW_IT; ++it) {
                const int c = c00 + it * 32 + lane;
                lc[it] = (c < nc) ? __ldg(hp + 4 + c) : -INFINITY;
            }
            float s[ROW_IT];
            unsigned pm[ROW_IT];
            bool any = false;
#pragma unroll
            for (int it = 0; it < ROW_IT; ++it) {
                s[it] = 0.f;
                bool p = false;
                if (lc[it] >= logit_floor) {
                    s[it] = __fsqrt_rn(__fmul_rn(sigmoid_f(lc[it]), o));
                    p = (s[it] >= thr);
                }
                pm[it] = __ballot_sync(0xffffffffu, p);
                any |= pm[it] != 0;
            }
            if (!any) continue;
            if (!size_known) {                                   // warp-uniform: every lane evaluates the same box
                cbox = decode_box(src.dec, hp, sc, cell, a);
                if (src.filter_small) size_ok = (__fsub_rn(cbox.z, cbox.x) > src.min_size) && (__fsub_rn(cbox.w, cbox.y) > src.min_size);
                size_known = true;
            }
            if (!size_ok) break;                                 // the whole row is dropped by filter_small_boxes
            int total = 0;
#pragma unroll
            for (int it = 0; it < ROW_IT; ++it) total += __popc(pm[it]);
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(O.counter, (unsigned long long)total);
            base = __shfl_sync(0xffffffffu, base, 0);
            int before = 0;
#pragma unroll
            for (int it = 0; it < ROW_IT; ++it) {
                if ((pm[it] >> lane) & 1u) {
                    const uint32_t seg = img * (uint32_t)nc + (uint32_t)(c00 + it * 32 + lane);
                    const int64_t pos = (int64_t)base + before + __popc(pm[it] & ((1u << lane) - 1u));
                    uint32_t sl = 0;
                    if (O.seg_cnt) sl = (uint32_t)atomicAdd(O.seg_cnt + seg, 1);      // classes of one row: all segments differ
                    if (pos < O.cap) {
                        O.keys[pos] = make_key(kl, seg, s[it], row);
                        O.vals[pos] = grow;
                        if (O.slot) O.slot[pos] = sl;
                        if (O.box) O.box[pos] = cbox;
                    }
                }
                before += __popc(pm[it]);
            }
        }
    }
}

__global__ void __launch_bounds__(256)
k_gather_sorted(CandSource src, const uint32_t* __restrict__ vals, int64_t n, float4* __restrict__ sbox,
                float* __restrict__ sarea, uint8_t* __restrict__ supp, uint8_t* __restrict__ keepf) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    float4 bx;
    if (src.from_heads) {
        const int64_t grow = vals[p];
        const int64_t img = grow / src.rows_per_image;
        int sc, cell, a;
        const float* ho;
        const float* hp = head_row(src.dec, (int)img, (int)(grow - img * src.rows_per_image), &sc, &cell, &a, &ho);
        bx = decode_box(src.dec, hp, sc, cell, a);
    } else {
        const float* b = src.box + (int64_t)vals[p] * src.box_stride;
        bx = make_float4(__ldg(b), __ldg(b + 1), __ldg(b + 2), __ldg(b + 3));
    }
    sbox[p] = bx;
    sarea[p] = box_area_exact(bx);
    supp[p] = 0;
    keepf[p] = 0;
}

__global__ void k_seg_offsets(const uint64_t* __restrict__ keys, int64_t n, int seg_shift, int nseg,
                              int64_t* __restrict__ off) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > nseg) return;
    int64_t lo = 0, hi = n;                     // first p with seg(key[p]) >= s
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if ((keys[mid] >> seg_shift) < (uint64_t)s) lo = mid + 1; else hi = mid;
    }
    off[s] = lo;
}

// ------------------------------------------------------------------------------------------
// Stable LSD radix sort of (u64 key, u32 value) pairs, 8 bits per pass, over the used key bits only.
//   k_rs_hist     per-block digit histogram (shared-memory atomics) -> hist[digit][block]
//   k_rs_scan     exclusive scan of the digit-major histogram (one block; it is 256 x nblocks ints)
//   k_rs_scatter  each block re-reads its elements IN ORDER, 256 at a time: rank inside the warp with
//                 __match_any_sync, prefix across the 8 warps per digit, running per-digit offsets
static constexpr int RS_THREADS = 256;
static constexpr int RS_ITEMS = 16;
static constexpr int RS_TILE = RS_THREADS * RS_ITEMS;

__global__ void __launch_bounds__(RS_THREADS)
k_rs_hist(const uint64_t* __restrict__ keys, int64_t n, int shift, int nblocks, int* __restrict__ hist) {
    __shared__ int s_h[256];
    s_h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const int64_t i = base + r * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&s_h[(int)((keys[i] >> shift) & 0xff)], 1);
    }
    __syncthreads();
    hist[(int64_t)threadIdx.x * nblocks + blockIdx.x] = s_h[threadIdx.x];
}

__global__ void __launch_bounds__(1024)
k_rs_scan(int* __restrict__ a, int64_t n) {
    __shared__ long long s_part[1024];
    const int64_t per = (n + 1023) / 1024;
    const int64_t b0 = (int64_t)threadIdx.x * per;
    long long sum = 0;
    for (int64_t i = 0; i < per; ++i) if (b0 + i < n) sum += a[b0 + i];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const long long v = (threadIdx.x >= o) ? s_part[threadIdx.x - o] : 0;
        __syncthreads();
        s_part[threadIdx.x] += v;
        __syncthreads();
    }
    long long run = s_part[threadIdx.x] - sum;
    for (int64_t i = 0; i < per; ++i)
        if (b0 + i < n) { const int c = a[b0 + i]; a[b0 + i] = (int)run; run += c; }
}

__global__ void __launch_bounds__(RS_THREADS)
k_rs_scatter(const uint64_t* __restrict__ kin, const uint32_t* __restrict__ vin, uint64_t* __restrict__ kout,
             uint32_t* __restrict__ vout, int64_t n, int shift, int nblocks, const int* __restrict__ hist, int scanned) {
    __shared__ int s_run[256];                 // next output slot of each digit for this block
    __shared__ int s_cnt[RS_THREADS / 32][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (scanned) {
        s_run[threadIdx.x] = hist[(int64_t)threadIdx.x * nblocks + blockIdx.x];
    } else {
        // raw per-block counts: thread d derives "keys with a smaller digit" + "same digit in earlier blocks" itself
        // (256 x nblocks ints, L2 resident) - saves the one-block scan launch of every pass
        const int* row = hist + (int64_t)threadIdx.x * nblocks;
        int total = 0, before = 0;
        for (int b = 0; b < nblocks; ++b) { const int c = row[b]; total += c; if (b < (int)blockIdx.x) before += c; }
        __shared__ int s_tot[256];
        s_tot[threadIdx.x] = total;
        __syncthreads();
        for (int o = 1; o < 256; o <<= 1) {
            const int v = (threadIdx.x >= o) ? s_tot[threadIdx.x - o] : 0;
            __syncthreads();
            s_tot[threadIdx.x] += v;
            __syncthreads();
        }
        s_run[threadIdx.x] = s_tot[threadIdx.x] - total + before;
    }
#pragma unroll
    for (int w = 0; w < RS_THREADS / 32; ++w) s_cnt[w][threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
    for (int r = 0; r < RS_ITEMS; ++r) {
        const int64_t i = base + r * RS_THREADS + threadIdx.x;
        const bool valid = i < n;
        uint64_t key = 0;
        uint32_t val = 0;
        int d = 256 + lane;                    // invalid lanes: unique pseudo-digits so they match nobody
        if (valid) { key = kin[i]; val = vin[i]; d = (int)((key >> shift) & 0xff); }
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank == 0) s_cnt[warp][d] = __popc(peers);
        __syncthreads();
        {   // thread t owns digit t: exclusive prefix over the warps (in order => stable), advance the run
            const int t = threadIdx.x;
            int run = s_run[t];
#pragma unroll
            for (int w = 0; w < RS_THREADS / 32; ++w) { const int c = s_cnt[w][t]; s_cnt[w][t] = run; run += c; }
            s_run[t] = run;
        }
        __syncthreads();
        if (valid) {
            const int pos = s_cnt[warp][d] + rank;
            kout[pos] = key;
            vout[pos] = val;
        }
        __syncthreads();
#pragma unroll
        for (int w = 0; w < RS_THREADS / 32; ++w) s_cnt[w][threadIdx.x] = 0;     // the prefix step overwrote every entry
        __syncthreads();
    }
}

// Segments of at most 256 boxes (the many-class regime: K2 has 5120 segments of ~50 boxes): one WARP per segment,
// PL boxes per lane (box j lives in slot j / 32 of lane j % 32), the alive set is PL 32-bit masks that every lane keeps
// in step through ballots.  Same greedy order and the same exact suppression test as the chunk path.
static constexpr int NMS_SMALL = 256;

template <int PL>
__device__ __forceinline__ void warp_nms(const float4* __restrict__ sbox, const float* __restrict__ sarea,
                                         uint8_t* __restrict__ keepf, int64_t s0, int m, float thr, int lane) {
    float4 b[PL];
    float a[PL];
    uint32_t alive[PL], kept[PL];
#pragma unroll
    for (int k = 0; k < PL; ++k) {
        const int j = k * 32 + lane;
        const bool valid = j < m;
        b[k] = valid ? sbox[s0 + j] : make_float4(0.f, 0.f, 0.f, 0.f);
        a[k] = valid ? sarea[s0 + j] : 0.f;
        alive[k] = __ballot_sync(0xffffffffu, valid);
        kept[k] = 0;
    }
#pragma unroll
    for (int k0 = 0; k0 < PL; ++k0) {
        while (alive[k0]) {
            const int src = __ffs((int)alive[k0]) - 1;
            alive[k0] &= ~(1u << src);
            kept[k0] |= 1u << src;
            float4 bi;
            bi.x = __shfl_sync(0xffffffffu, b[k0].x, src);
            bi.y = __shfl_sync(0xffffffffu, b[k0].y, src);
            bi.z = __shfl_sync(0xffffffffu, b[k0].z, src);
            bi.w = __shfl_sync(0xffffffffu, b[k0].w, src);
            const float ai = __shfl_sync(0xffffffffu, a[k0], src);
#pragma unroll
            for (int k = k0; k < PL; ++k) {
                const bool kill = ((alive[k] >> lane) & 1u) && suppresses_exact(bi, ai, b[k], a[k], thr);
                alive[k] &= ~__ballot_sync(0xffffffffu, kill);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < PL; ++k)
        if ((kept[k] >> lane) & 1u) keepf[s0 + k * 32 + lane] = 1;
}

__global__ void __launch_bounds__(256)
k_nms_small(const float4* __restrict__ sbox, const float* __restrict__ sarea, uint8_t* __restrict__ keepf,
            const int64_t* __restrict__ seg_off, int nseg, float thr) {
    const int lane = threadIdx.x & 31;
    const int warp_g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    for (int seg = warp_g; seg < nseg; seg += n_warps) {
        const int64_t s0 = seg_off[seg];
        const int m = (int)min((int64_t)(NMS_SMALL + 1), seg_off[seg + 1] - s0);
        if (m <= 0 || m > NMS_SMALL) continue;
        if (m <= 64) warp_nms<2>(sbox, sarea, keepf, s0, m, thr, lane);
        else if (m <= 128) warp_nms<4>(sbox, sarea, keepf, s0, m, thr, lane);
        else warp_nms<8>(sbox, sarea, keepf, s0, m, thr, lane);
    }
}

__global__ void __launch_bounds__(NMS_THREADS)
k_nms_segments(const float4* __restrict__ sbox, const float* __restrict__ sarea, uint8_t* __restrict__ supp,
               uint8_t* __restrict__ keepf, const int64_t* __restrict__ seg_off, int nseg, float thr,
               int64_t big_segment, int64_t small_segment) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ChunkSmem& S = *reinterpret_cast<ChunkSmem*>(smem_raw);
    __shared__ int s_idx[NMS_T];
    __shared__ int s_wcount[NMS_THREADS / 32];
    __shared__ int s_n;
    __shared__ long long s_pos;
    for (int seg = blockIdx.x; seg < nseg; seg += gridDim.x) {
        const int64_t s0 = seg_off[seg];
        const int64_t m = seg_off[seg + 1] - s0;
        if (m <= small_segment || m > big_segment) continue;       // small ones: k_nms_small
        // steps of <= NMS_T still-alive boxes (gathered behind a cursor, so boxes that earlier steps suppressed
        // cost nothing any more): resolve the step, then apply its kept boxes to the rest of the segment
        if (threadIdx.x == 0) s_pos = 0;
        __syncthreads();
        while (true) {
            const int ct = gather_alive(supp, s0, (long long)m, &s_pos, &s_n, s_idx, s_wcount);
            if (ct == 0) break;
            chunk_resolve(S, sbox, sarea, supp, s0, ct, thr, s_idx);
            const int nk = S.nk;
            for (int t = threadIdx.x; t < nk; t += NMS_THREADS) keepf[s0 + s_idx[S.klist[t]]] = 1;
            const long long first = s_pos;
            for (long long j = first + threadIdx.x; j < m; j += NMS_THREADS) {
                if (supp[s0 + j]) continue;
                const float4 bj = sbox[s0 + j];
                const float aj = sarea[s0 + j];
                for (int t = 0; t < nk; ++t) {
                    const int k = S.klist[t];
                    if (suppresses_exact(S.box[k], S.area[k], bj, aj, thr)) { supp[s0 + j] = 1; break; }
                }
            }
            __syncthreads();
        }
    }
}

// big-segment path: resolve one chunk on one CTA, publish its kept boxes...
__global__ void __launch_bounds__(NMS_THREADS)
k_nms_resolve(const float4* __restrict__ sbox, const float* __restrict__ sarea, const uint8_t* __restrict__ supp,
              uint8_t* __restrict__ keepf, int64_t base, int ct, float thr, float4* __restrict__ kbox,
              float* __restrict__ karea, int* __restrict__ knum) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ChunkSmem& S = *reinterpret_cast<ChunkSmem*>(smem_raw);
    chunk_resolve(S, sbox, sarea, supp, base, ct, thr);
    const int nk = S.nk;
    for (int t = threadIdx.x; t < nk; t += NMS_THREADS) {
        const int k = S.klist[t];
        keepf[base + k] = 1;
        kbox[t] = S.box[k];
        karea[t] = S.area[k];
    }
    if (threadIdx.x == 0) *knum = nk;
}

// ...and apply them to every later box of the segment with the whole GPU.
__global__ void __launch_bounds__(256)
k_nms_apply(const float4* __restrict__ sbox, const float* __restrict__ sarea, uint8_t* __restrict__ supp,
            int64_t first, int64_t last, float thr, const float4* __restrict__ kbox,
            const float* __restrict__ karea, const int* __restrict__ knum) {
    __shared__ float4 s_box[NMS_T];
    __shared__ float s_area[NMS_T];
    const int nk = *knum;
    for (int t = threadIdx.x; t < nk; t += blockDim.x) { s_box[t] = kbox[t]; s_area[t] = karea[t]; }
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = first + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < last; j += stride) {
        if (supp[j]) continue;
        const float4 bj = sbox[j];
        const float aj = sarea[j];
        for (int t = 0; t < nk; ++t) {
            if (suppresses_exact(s_box[t], s_area[t], bj, aj, thr)) { supp[j] = 1; break; }
        }
    }
}

// Very large segments, cursor form: instead of fixed 512-box windows (most of whose boxes are long dead deep into
// the segment), every step gathers the NEXT 512 still-alive boxes after a device-side cursor, resolves them, and
// applies their kept boxes to everything after the new cursor.  Same greedy order, ~5x fewer steps at 200 k boxes.
struct BigState {
    long long cursor;      // first segment-relative position not yet gathered
    int knum;              // kept boxes published by the last resolve
    int done;              // cursor reached the end of the segment
};

__global__ void __launch_bounds__(NMS_THREADS)
k_nms_resolve_next(const float4* __restrict__ sbox, const float* __restrict__ sarea, const uint8_t* __restrict__ supp,
                   uint8_t* __restrict__ keepf, int64_t s0, int64_t m, float thr, float4* __restrict__ kbox,
                   float* __restrict__ karea, BigState* __restrict__ st) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ChunkSmem& S = *reinterpret_cast<ChunkSmem*>(smem_raw);
    __shared__ int s_idx[NMS_T];
    __shared__ int s_wcount[NMS_THREADS / 32];
    __shared__ int s_n;
    __shared__ long long s_pos;
    if (threadIdx.x == 0) s_pos = st->cursor;
    __syncthreads();
    if (s_pos >= m) {
        if (threadIdx.x == 0) { st->knum = 0; st->done = 1; }
        return;
    }
    gather_alive(supp, s0, (long long)m, &s_pos, &s_n, s_idx, s_wcount);
    const int ct = s_n;
    if (ct > 0) chunk_resolve(S, sbox, sarea, supp, s0, ct, thr, s_idx);
    const int nk = ct > 0 ? S.nk : 0;
    for (int t = threadIdx.x; t < nk; t += NMS_THREADS) {
        const int k = S.klist[t];
        keepf[s0 + s_idx[k]] = 1;
        kbox[t] = S.box[k];
        karea[t] = S.area[k];
    }
    if (threadIdx.x == 0) { st->knum = nk; st->cursor = s_pos; st->done = (s_pos >= m) ? 1 : 0; }
}

__global__ void __launch_bounds__(256)
k_nms_apply_from(const float4* __restrict__ sbox, const float* __restrict__ sarea, uint8_t* __restrict__ supp,
                 int64_t s0, int64_t m, float thr, const float4* __restrict__ kbox, const float* __restrict__ karea,
                 const BigState* __restrict__ st) {
    __shared__ float4 s_box[NMS_T];
    __shared__ float s_area[NMS_T];
    const int nk = st->knum;
    const long long first = st->cursor;
    if (nk == 0 || first >= m) return;
    for (int t = threadIdx.x; t < nk; t += blockDim.x) { s_box[t] = kbox[t]; s_area[t] = karea[t]; }
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = s0 + first + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < s0 + m; j += stride) {
        if (supp[j]) continue;
        const float4 bj = sbox[j];
        const float aj = sarea[j];
        for (int t = 0; t < nk; ++t) {
            if (suppresses_exact(s_box[t], s_area[t], bj, aj, thr)) { supp[j] = 1; break; }
        }
    }
}

// ------------------------------------------------------------------------------------------
// ordered compaction of keep flags
__global__ void __launch_bounds__(CMP_BLOCK)
k_count(const uint8_t* __restrict__ flags, int64_t n, int* __restrict__ blk) {
    __shared__ int s_w[CMP_BLOCK / 32];
    const int64_t p = (int64_t)blockIdx.x * CMP_BLOCK + threadIdx.x;
    const bool f = p < n && flags[p];
    const unsigned b = __ballot_sync(0xffffffffu, f);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = __popc(b);
    __syncthreads();
    if (threadIdx.x < 32) {
        int v = s_w[threadIdx.x];
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) blk[blockIdx.x] = v;
    }
}

// single block exclusive scan of nb block counts; total -> *total_out
__global__ void __launch_bounds__(1024)
k_scan(int* __restrict__ blk, int nb, long long* __restrict__ total_out) {
    __shared__ long long s_part[1024];
    const int per = (nb + 1023) / 1024;
    const int b0 = threadIdx.x * per;
    long long sum = 0;
    for (int i = 0; i < per; ++i) if (b0 + i < nb) sum += blk[b0 + i];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {            // Hillis-Steele inclusive scan
        long long v = (threadIdx.x >= o) ? s_part[threadIdx.x - o] : 0;
        __syncthreads();
        s_part[threadIdx.x] += v;
        __syncthreads();
    }
    long long run = s_part[threadIdx.x] - sum;      // exclusive prefix of this thread's span
    for (int i = 0; i < per; ++i) {
        if (b0 + i < nb) { const int c = blk[b0 + i]; blk[b0 + i] = (int)run; run += c; }
    }
    if (threadIdx.x == 1023) *total_out = s_part[1023];
}

__global__ void __launch_bounds__(CMP_BLOCK)
k_scatter(const uint8_t* __restrict__ flags, int64_t n, const int* __restrict__ blk,
          const uint64_t* __restrict__ keys, const float4* __restrict__ sbox, KeyLayout kl, int nc,
          float4* __restrict__ o_box, float* __restrict__ o_score, int32_t* __restrict__ o_label,
          int32_t* __restrict__ o_img, int32_t* __restrict__ o_src) {
    __shared__ int s_w[CMP_BLOCK / 32];
    const int64_t p = (int64_t)blockIdx.x * CMP_BLOCK + threadIdx.x;
    const bool f = p < n && flags[p];
    const int rank = block_rank(f, s_w);
    if (f) {
        const int64_t q = (int64_t)blk[blockIdx.x] + rank;
        const uint64_t key = keys[p];
        const uint64_t seg = key >> kl.seg_shift;
        o_box[q] = sbox[p];
        o_score[q] = from_orderable(~(((uint32_t)(key >> kl.row_bits) & kl.score_mask) + kl.score_base));
        o_label[q] = (int32_t)(seg % (uint64_t)nc);
        o_img[q] = (int32_t)(seg / (uint64_t)nc);
        o_src[q] = (int32_t)(key & kl.row_mask);
    }
}

// ------------------------------------------------------------------------------------------
int64_t PostProc::flag_offsets(const uint8_t* flags, int64_t n) {
    cudaStream_t st = ctx->stream;
    if (n <= 0) return 0;
    const int nb = ceil_div(n, CMP_BLOCK);
    blk.reserve((size_t)nb * 4);
    counters.reserve(64);
    host_small.reserve(64);
    k_count<<<nb, CMP_BLOCK, 0, st>>>(flags, n, blk.as<int>());
    Y3_LAUNCHED(ctx);
    k_scan<<<1, 1024, 0, st>>>(blk.as<int>(), nb, reinterpret_cast<long long*>(counters.as<unsigned char>() + 8));
    Y3_LAUNCHED(ctx);
    unsigned long long* h = host_small.as<unsigned long long>();
    Y3_CUDA(cudaMemcpyAsync(h + 1, counters.as<unsigned char>() + 8, 8, cudaMemcpyDeviceToHost, st));
    Y3_CUDA(cudaStreamSynchronize(st));
    return (int64_t)h[1];
}

KeyLayout PostProc::key_layout(const CandSource& src) const {
    const int64_t nseg64 = src.num_segments();
    KeyLayout kl;
    kl.row_bits = ilog2_ceil((uint64_t)src.rows_per_image);
    if (kl.row_bits == 0) kl.row_bits = 1;
    int score_bits = 32;
    kl.score_base = 0;
    if (src.from_heads && src.score_thr > 0.f && src.score_thr < 1.f) {
        // host mirror of orderable(): positive floats map to bits | 0x80000000
        uint32_t hi_bits, lo_bits;
        const float one = 1.0f, thr = src.score_thr;
        memcpy(&hi_bits, &one, 4); memcpy(&lo_bits, &thr, 4);
        kl.score_base = ~(hi_bits | 0x80000000u);                       // ~orderable(1.0f): the smallest field value
        const uint32_t span = ~(lo_bits | 0x80000000u) - kl.score_base; // ~orderable(thr) - base: the largest
        score_bits = ilog2_ceil((uint64_t)span + 1);
        if (score_bits < 1) score_bits = 1;
    }
    kl.score_mask = score_bits >= 32 ? 0xffffffffu : ((1u << score_bits) - 1u);
    kl.seg_shift = kl.row_bits + score_bits;
    const int seg_bits = ilog2_ceil((uint64_t)nseg64 + 1);
    kl.total_bits = kl.seg_shift + seg_bits;
    kl.row_mask = (1ull << kl.row_bits) - 1ull;
    Y3_CHECK(kl.total_bits <= 64, Y3_ERR_UNSUPPORTED, "sort key needs %d bits", kl.total_bits);
    return kl;
}

// Capacity of the candidate list: y3_config.max_candidates, else every (row, class) pair up to 4 Mi entries and never
// fewer than the rows (so that a single class can always pass completely).
int64_t PostProc::capacity(const CandSource& src) const {
    const int64_t rows = src.rows_per_image * src.n_images;
    const int64_t total = rows * src.nc;
    int64_t cap = ctx->cfg.max_candidates > 0 ? ctx->cfg.max_candidates : std::min<int64_t>(total, std::max<int64_t>(rows, 4ll << 20));
    return std::min(cap, total);
}

// threshold + compaction (+ fused decode) into keys[0] / vals[0]; counters[0] = candidate count (zeroed here)
void PostProc::launch_candidates(const CandSource& src, const KeyLayout& kl, int64_t cap, bool count_segments) {
    cudaStream_t st = ctx->stream;
    const int64_t rows = src.rows_per_image * src.n_images;
    const int64_t total = rows * src.nc;
    keys[0].reserve(cap * 8);
    vals[0].reserve(cap * 4);
    counters.reserve(64);
    Y3_CUDA(cudaMemsetAsync(counters.p, 0, 64, st));
    CandOut O{keys[0].as<uint64_t>(), vals[0].as<uint32_t>(), counters.as<unsigned long long>(), cap, nullptr, nullptr, nullptr};
    if (count_segments) {
        slot.reserve(cap * 4);
        cbox.reserve(cap * 16);
        O.seg_cnt = seg_cnt.as<int>();
        O.slot = slot.as<uint32_t>();
        O.box = cbox.as<float4>();
    }
    // logit(thr^2) minus a margin far above any rounding of expf / the division (see k_candidates)
    float logit_floor = -INFINITY;
    if (src.from_heads && src.score_thr > 0.f && src.score_thr < 1.f) {
        const double t2 = (double)src.score_thr * (double)src.score_thr;
        logit_floor = (float)(log(t2 / (1.0 - t2)) - 0.05);
    }
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)ctx->sm_count * 16);
    static const bool row_kernel = getenv("Y3_CAND_OLD") == nullptr;
    if (row_kernel && src.from_heads && total + 32 < (1ll << 32)) {
        const int rblocks = (int)std::min<int64_t>((rows + 255) / 256, (int64_t)ctx->sm_count * 8);
        if (src.nc == 1) {
            launch_chained(ctx, k_candidates_heads1, rblocks, 256, 0, st, src, kl, logit_floor, O);
            return;
        } else {
            live.reserve((size_t)rows * 4);
            unsigned int* n_live = reinterpret_cast<unsigned int*>(counters.as<unsigned char>() + 16);     // zeroed above
            launch_chained(ctx, k_live_rows, rblocks, 256, 0, st, src, logit_floor, live.as<uint32_t>(), n_live);
            launch_chained(ctx, k_candidates_rows, ctx->sm_count * 8, 256, 0, st, src, kl, logit_floor, (const uint32_t*)live.as<uint32_t>(),
                           (const unsigned int*)n_live, O);
            return;
        }
    } else if (total + 32 < (1ll << 32)) {
        k_candidates<uint32_t><<<blocks, 256, 0, st>>>(src, kl, total, logit_floor, O);
    } else {
        k_candidates<uint64_t><<<blocks, 256, 0, st>>>(src, kl, total, logit_floor, O);
    }
    Y3_LAUNCHED(ctx);
}

void PostProc::enqueue(const CandSource& src, float iou_thr) {
    cudaStream_t st = ctx->stream;
    pending = NmsResult();
    pending_cap = -1;
    const int64_t rows = src.rows_per_image * src.n_images;
    const int64_t total = rows * src.nc;
    if (total <= 0) return;
    Y3_CHECK(rows < (1ll << 32), Y3_ERR_UNSUPPORTED, "too many rows (%lld)", (long long)rows);
    const KeyLayout kl = key_layout(src);
    const int64_t cap = capacity(src);
    static const bool no_seg = getenv("Y3_NMS_GLOBAL_SORT") != nullptr;      // A/B: the round-1 global-sort pipeline only
    if (!no_seg && src.num_segments() < (1ll << 24) && kl.seg_shift <= 62) {
        // segmented pipeline: bin by (image, class), per-segment NMS.  Segments are bounded by rows_per_image; when that
        // bound exceeds what one CTA sorts in shared memory the largest segment is read back before the route is chosen.
        segmented_front(src, kl, cap);
        if (segmented_ok(src)) {
            segmented_nms(src, kl, iou_thr, nullptr);
            segmented_emit_plain(src, kl);
            return;
        }
        host_ctrl.reserve(sizeof(PostCtrl));
        Y3_CUDA(cudaMemcpyAsync(host_ctrl.p, ctrl.p, sizeof(PostCtrl), cudaMemcpyDeviceToHost, st));
        Y3_CUDA(cudaStreamSynchronize(st));
        const PostCtrl C = *host_ctrl.as<PostCtrl>();
        Y3_CHECK(!C.overflow, Y3_ERR_NOSPACE, "candidate list overflow: %llu candidates, capacity %lld (raise y3_config.max_candidates)",
                 C.n_cand, (long long)cap);
        pending.n_cand = (int64_t)C.n_cand;
        if (C.K == 0) return;
        if (C.max_seg <= SEG_MID_MAX) {
            segmented_nms(src, kl, iou_thr, nullptr);
            segmented_emit_plain(src, kl);
            return;
        }
        pending = run_global_sort(src, iou_thr, kl, cap, C.K);           // keys[0] / vals[0] hold the unsorted list
        return;
    }
    if (!ev0) { Y3_CUDA(cudaEventCreate(&ev0)); Y3_CUDA(cudaEventCreate(&ev1)); }
    Y3_CUDA(cudaEventRecord(ev0, st));
    launch_candidates(src, kl, cap, false);
    Y3_CUDA(cudaEventRecord(ev1, st));
    host_small.reserve(64);
    unsigned long long* h_cnt = host_small.as<unsigned long long>();
    Y3_CUDA(cudaMemcpyAsync(h_cnt, counters.p, 8, cudaMemcpyDeviceToHost, st));
    Y3_CUDA(cudaStreamSynchronize(st));
    const int64_t K = (int64_t)h_cnt[0];
    { float ms = 0.f; if (cudaEventElapsedTime(&ms, ev0, ev1) == cudaSuccess) ctx->timings.ms_decode += ms; }
    Y3_CHECK(K <= cap, Y3_ERR_NOSPACE, "candidate list overflow: %lld candidates, capacity %lld "
             "(raise y3_config.max_candidates)", (long long)K, (long long)cap);
    pending.n_cand = K;
    if (K == 0) return;
    pending = run_global_sort(src, iou_thr, kl, cap, K);
}

// The round-1 pipeline from the unsorted candidate list on: global radix sort by (segment, score desc, row asc),
// segment offsets, NMS (whole-GPU apply phase for very large segments), ordered compaction.  Used when a segment can
// exceed what one CTA sorts in shared memory (e.g. single_class_nms on 200 k boxes).  Synchronises the stream.
NmsResult PostProc::run_global_sort(const CandSource& src, float iou_thr, const KeyLayout& kl, int64_t cap, int64_t K) {
    cudaStream_t st = ctx->stream;
    NmsResult R;
    R.n_cand = K;
    const int nseg = (int)src.num_segments();
    keys[1].reserve(cap * 8);
    vals[1].reserve(cap * 4);
    host_small.reserve(64 + (size_t)(nseg + 1) * 8);

    // sort by (segment, score desc, row asc): own stable LSD radix sort over the used key bits
    Y3_CHECK(K < (1ll << 31), Y3_ERR_UNSUPPORTED, "too many candidates (%lld)", (long long)K);
    int src_buf = 0;
    {
        const int nblocks = ceil_div(K, RS_TILE);
        sort_tmp.reserve((size_t)256 * nblocks * 4);
        for (int shift = 0; shift < kl.total_bits; shift += 8) {
            const int dst_buf = src_buf ^ 1;
            k_rs_hist<<<nblocks, RS_THREADS, 0, st>>>(keys[src_buf].as<uint64_t>(), K, shift, nblocks, sort_tmp.as<int>());
            Y3_LAUNCHED(ctx);
            const int scanned = nblocks > 512 ? 1 : 0;     // few blocks: every scatter block sums the histogram itself
            if (scanned) {
                k_rs_scan<<<1, 1024, 0, st>>>(sort_tmp.as<int>(), (int64_t)256 * nblocks);
                Y3_LAUNCHED(ctx);
            }
            k_rs_scatter<<<nblocks, RS_THREADS, 0, st>>>(keys[src_buf].as<uint64_t>(), vals[src_buf].as<uint32_t>(),
                                                        keys[dst_buf].as<uint64_t>(), vals[dst_buf].as<uint32_t>(), K, shift, nblocks,
                                                        sort_tmp.as<int>(), scanned);
            Y3_LAUNCHED(ctx);
            src_buf = dst_buf;
        }
    }
    const uint64_t* skeys = keys[src_buf].as<uint64_t>();
    const uint32_t* svals = vals[src_buf].as<uint32_t>();

    sbox.reserve(K * 16); sarea.reserve(K * 4); supp.reserve(K); keepf.reserve(K);
    seg_off.reserve((size_t)(nseg + 1) * 8);
    k_gather_sorted<<<ceil_div(K, 256), 256, 0, st>>>(src, svals, K, sbox.as<float4>(), sarea.as<float>(),
                                                      supp.as<uint8_t>(), keepf.as<uint8_t>());
    Y3_LAUNCHED(ctx);
    k_seg_offsets<<<ceil_div(nseg + 1, 256), 256, 0, st>>>(skeys, K, kl.seg_shift, nseg, seg_off.as<int64_t>());
    Y3_LAUNCHED(ctx);

    // segment sizes are needed on the host only to route very large segments
    int64_t* h_off = reinterpret_cast<int64_t*>(host_small.as<unsigned char>() + 64);
    bool any_big = false, any_small = false;
    if (K > BIG_SEGMENT && src.rows_per_image > BIG_SEGMENT) {      // a segment never exceeds rows_per_image
        Y3_CUDA(cudaMemcpyAsync(h_off, seg_off.p, (size_t)(nseg + 1) * 8, cudaMemcpyDeviceToHost, st));
        Y3_CUDA(cudaStreamSynchronize(st));
        for (int s = 0; s < nseg; ++s) {
            const int64_t m = h_off[s + 1] - h_off[s];
            if (m > BIG_SEGMENT) any_big = true; else if (m > 0) any_small = true;
        }
    } else {
        any_small = true;
    }

    const size_t smem = sizeof(ChunkSmem);
    static_assert(sizeof(ChunkSmem) <= 48 * 1024, "ChunkSmem must fit the default dynamic smem limit");
    if (any_small) {
        static const bool warp_nms = getenv("Y3_NO_WARP_NMS") == nullptr;
        if (warp_nms && nseg > 1) {
            const int wblocks = std::min((nseg + 7) / 8, ctx->sm_count * 6);
            k_nms_small<<<wblocks, 256, 0, st>>>(sbox.as<float4>(), sarea.as<float>(), keepf.as<uint8_t>(), seg_off.as<int64_t>(), nseg,
                                                  iou_thr);
            Y3_LAUNCHED(ctx);
        }
        const int blocks = std::min(nseg, ctx->sm_count * 4);
        k_nms_segments<<<blocks, NMS_THREADS, smem, st>>>(sbox.as<float4>(), sarea.as<float>(), supp.as<uint8_t>(),
                                                          keepf.as<uint8_t>(), seg_off.as<int64_t>(), nseg, iou_thr,
                                                          BIG_SEGMENT, (warp_nms && nseg > 1) ? (int64_t)NMS_SMALL : 0);
        Y3_LAUNCHED(ctx);
    }
    if (any_big) {
        kbuf.reserve((size_t)NMS_T * 20 + 16 + 64);
        float4* kbox = kbuf.as<float4>();
        float* karea = reinterpret_cast<float*>(kbuf.as<unsigned char>() + (size_t)NMS_T * 16);
        int* knum = reinterpret_cast<int*>(kbuf.as<unsigned char>() + (size_t)NMS_T * 20);
        for (int s = 0; s < nseg; ++s) {
            const int64_t s0 = h_off[s], m = h_off[s + 1] - h_off[s];
            if (m <= BIG_SEGMENT) continue;
            static const bool cursor_form = getenv("Y3_NMS_FIXED_CHUNKS") == nullptr;
            if (cursor_form) {
                BigState* bs = reinterpret_cast<BigState*>(kbuf.as<unsigned char>() + (size_t)NMS_T * 20 + 16);
                Y3_CUDA(cudaMemsetAsync(bs, 0, sizeof(BigState), st));
                BigState* h_bs = reinterpret_cast<BigState*>(host_small.as<unsigned char>() + 32);
                const int blocks = (int)std::min<int64_t>((m + 255) / 256, (int64_t)ctx->sm_count * 8);
                const int64_t max_steps = (m + NMS_T - 1) / NMS_T;          // every step consumes >= NMS_T positions or ends
                for (int64_t step = 0; step < max_steps;) {
                    const int burst = (int)std::min<int64_t>(16, max_steps - step);
                    for (int b = 0; b < burst; ++b) {
                        k_nms_resolve_next<<<1, NMS_THREADS, smem, st>>>(sbox.as<float4>(), sarea.as<float>(), supp.as<uint8_t>(),
                                                                         keepf.as<uint8_t>(), s0, m, iou_thr, kbox, karea, bs);
                        Y3_LAUNCHED(ctx);
                        k_nms_apply_from<<<blocks, 256, 0, st>>>(sbox.as<float4>(), sarea.as<float>(), supp.as<uint8_t>(), s0, m,
                                                                 iou_thr, kbox, karea, bs);
                        Y3_LAUNCHED(ctx);
                    }
                    step += burst;
                    Y3_CUDA(cudaMemcpyAsync(h_bs, bs, sizeof(BigState), cudaMemcpyDeviceToHost, st));
                    Y3_CUDA(cudaStreamSynchronize(st));
                    if (h_bs->done) break;
                }
                continue;
            }
            for (int64_t c0 = 0; c0 < m; c0 += NMS_T) {
                const int ct = (int)std::min<int64_t>(NMS_T, m - c0);
                k_nms_resolve<<<1, NMS_THREADS, smem, st>>>(sbox.as<float4>(), sarea.as<float>(), supp.as<uint8_t>(),
                                                            keepf.as<uint8_t>(), s0 + c0, ct, iou_thr, kbox, karea, knum);
                Y3_LAUNCHED(ctx);
                const int64_t first = s0 + c0 + ct, last = s0 + m;
                if (first < last) {
                    const int blocks = (int)std::min<int64_t>((last - first + 255) / 256, (int64_t)ctx->sm_count * 8);
                    k_nms_apply<<<blocks, 256, 0, st>>>(sbox.as<float4>(), sarea.as<float>(), supp.as<uint8_t>(),
                                                        first, last, iou_thr, kbox, karea, knum);
                    Y3_LAUNCHED(ctx);
                }
            }
        }
    }

    // ordered compaction of the kept candidates
    const int nb = ceil_div(K, CMP_BLOCK);
    const int64_t kept = flag_offsets(keepf.as<uint8_t>(), K);
    R.n_kept = kept;
    if (kept == 0) return R;
    o_box.reserve(kept * 16); o_score.reserve(kept * 4); o_label.reserve(kept * 4);
    o_img.reserve(kept * 4); o_src.reserve(kept * 4);
    k_scatter<<<nb, CMP_BLOCK, 0, st>>>(keepf.as<uint8_t>(), K, blk.as<int>(), skeys, sbox.as<float4>(), kl, src.nc,
                                        o_box.as<float4>(), o_score.as<float>(), o_label.as<int32_t>(),
                                        o_img.as<int32_t>(), o_src.as<int32_t>());
    Y3_LAUNCHED(ctx);
    R.boxes = o_box.as<float4>(); R.scores = o_score.as<float>(); R.labels = o_label.as<int32_t>();
    R.img = o_img.as<int32_t>(); R.src_row = o_src.as<int32_t>();
    return R;
}

}  // namespace y3
