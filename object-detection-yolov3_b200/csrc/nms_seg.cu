// nms_seg.cu - the synchronisation-free, segmented post-processing pipeline (round 2).
//
// Same results, bit for bit, as the global-sort pipeline of nms.cu (reference: bbox_utils.py:217-281 =
// inference_tiled.py:120-182; ordering rule score descending, then row ascending), but organised around the
// (image, class) SEGMENT instead of one global sort, and without a single host round trip between the kernels:
//
//   k_candidates*   (nms.cu) threshold + compaction; every candidate also takes the next slot of its segment
//   k2_scan         one CTA: exclusive scan of the per-segment counts -> segment offsets, ONE work list of the
//                   non-empty segments ordered by size class (largest first), overflow / largest-segment flags - all
//                   into a device control block
//   k2_bin          scatters the 64-bit keys and boxes into their segment's range (arrival order inside a segment)
//                   (k2_scan_bin: both in one launch when the segment table fits shared memory - every block scans
//                   the counts itself, block 0 publishes the global results)
//   k2_nms_warp     segments of <= 128 boxes, one WARP each: no sort at all - greedy NMS by selection: the alive box
//                   with the smallest key (= highest score, lowest row) is found with two warp min-reductions (REDUX),
//                   emitted, and the boxes it suppresses are cleared; picks come out in the reference's output order
//   k2_nms_cta      segments of 129 .. 24576 boxes, one CTA each (device-side cursor over the work list), on an
//                   auxiliary stream next to the warp kernel: bitonic sort of the keys in shared memory, boxes gathered
//                   in sorted order, the chunked bitmask NMS of nms.cu, ordered compaction
//   k2_out_scan     one CTA: exclusive scan of the per-segment kept counts -> output offsets, totals
//   k2_emit         one warp per segment copies its kept records to the outputs.  On the tiled path the seam
//                   stitching (inference_tiled.py:235-301) is folded in: the ownership test runs where the kept box
//                   is produced, and this kernel writes the final float64 [x0,y0,x1,y1,score,label] rows.
//                   (k2_emit_fused: scan + emit in one launch for small segment tables)
//
// Every count (candidates, segment sizes, kept boxes, accumulated rows) stays on the device; launches use fixed grids
// that read their bounds from the control block.  The plain entry points read the kept count back once, the tiled
// path reads the control block once per call (after the last tile batch).
#include "nms_common.cuh"

#include <algorithm>
#include <stdlib.h>

namespace y3 {

// ------------------------------------------------------------------------------------------ scan / bin
// exclusive prefix of one int per thread over a 1024-thread block (warp shuffles + one shared round); *total = block sum
__device__ __forceinline__ int block_excl_scan_1024(int v, int* s_warp /*[33]*/, int* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        const int w = s_warp[lane];
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        s_warp[lane] = wi - w;
        if (lane == 31) s_warp[32] = wi;
    }
    __syncthreads();
    *total = s_warp[32];
    return s_warp[wid] + incl - v;
}

// Size class of a non-empty segment, largest first:
//   0..2  CTA-resolved with the large key buffer  (> 16384, > 8192, > 4096 boxes)
//   3..7  CTA-resolved, keys fit the shared memory the NMS phase needs anyway (> 2048, > 1024, > 512, > 256, > 128)
//   8..10 warp-resolved (> 64, > 32, <= 32)
// The work list is ordered by class: the CTA queues start with their longest segments, and the static round-robin
// of the warp kernel deals every warp a similar mix.
static constexpr int N_CLS = 11, N_CTAB_CLS = 3, N_CTA_CLS = 8;
static constexpr int SEG_CTAA_MAX = 4096;
__device__ __forceinline__ int size_class(int c) {
    if (c > SEG_WARP_MAX) {
        int k = N_CTA_CLS - 1;
        for (int lim = 2 * SEG_WARP_MAX; c > lim && k > 0; lim <<= 1) --k;
        return k;
    }
    return c > 64 ? 8 : c > 32 ? 9 : 10;
}

// Exclusive scan of the per-segment counts + work list + control block, by ONE block of 1024 threads.
//   s_table (optional, shared memory, nseg + 1 ints): the offsets also land there (k2_scan_bin: every block keeps its own copy)
//   publish: this block writes the global results (segment offsets, work list, kept-count zeros, control block)
__device__ __forceinline__ void seg_scan_block(const int* __restrict__ seg_cnt, int nseg, int* __restrict__ seg_off, int* __restrict__ work_list,
                                               int* __restrict__ kept_cnt, const unsigned long long* __restrict__ counter, long long cap,
                                               PostCtrl* __restrict__ C, int* s_table, bool publish, int* total_out, bool* overflow_out) {
    __shared__ int s_warp[33];
    __shared__ int s_cls[N_CLS], s_cls_off[N_CLS], s_max;
    const int tid = threadIdx.x;
    const unsigned long long n_cand = *counter;
    const bool overflow = n_cand > (unsigned long long)cap;
    const int per = (nseg + 1023) / 1024;
    const int b0 = tid * per;
    if (tid < N_CLS) s_cls[tid] = 0;
    if (tid == 0) s_max = 0;
    __syncthreads();
    int sum = 0, mx = 0;
    for (int i = 0; i < per; ++i)
        if (b0 + i < nseg) {
            const int c = seg_cnt[b0 + i];
            sum += c; mx = max(mx, c);
            if (c >= 1 && c <= SEG_MID_MAX) atomicAdd(&s_cls[size_class(c)], 1);
        }
    for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((tid & 31) == 0 && mx) atomicMax(&s_max, mx);
    int total = 0;
    int run = block_excl_scan_1024(sum, s_warp, &total);        // (contains the __syncthreads that publish s_cls / s_max)
    if (tid == 0) {
        int acc = 0;
        for (int k = 0; k < N_CLS; ++k) { s_cls_off[k] = acc; acc += s_cls[k]; }
    }
    __syncthreads();
    for (int i = 0; i < per; ++i) {
        const int s = b0 + i;
        if (s >= nseg) break;
        const int c = seg_cnt[s];
        if (s_table) s_table[s] = overflow ? 0 : run;
        if (publish) {
            seg_off[s] = overflow ? 0 : run;
            if (overflow || c == 0 || c > SEG_MID_MAX) kept_cnt[s] = 0;          // nobody else will write it
            else work_list[atomicAdd(&s_cls_off[size_class(c)], 1)] = s;
        }
        run += c;
    }
    __syncthreads();
    *total_out = overflow ? 0 : total;
    *overflow_out = overflow;
    if (tid == 0 && publish) {
        int n_b = 0, n_cta = 0, n_all = 0;
        for (int k = 0; k < N_CLS; ++k) { n_all += s_cls[k]; if (k < N_CTA_CLS) n_cta += s_cls[k]; if (k < N_CTAB_CLS) n_b += s_cls[k]; }
        seg_off[nseg] = overflow ? 0 : total;
        C->n_cand = n_cand;
        C->overflow = overflow ? 1 : 0;
        C->K = overflow ? 0 : total;
        C->n_big = overflow ? 0 : n_b;                           // work_list[0, n_big): CTA-resolved, large key buffer
        C->big_next = 0;
        C->n_mid = overflow ? 0 : n_cta;                         // work_list[n_big, n_mid): CTA-resolved
        C->mid_next = C->n_big;
        C->n_small = overflow ? 0 : n_all;                       // work_list[n_mid, n_small): warp-resolved
        C->max_seg = s_max;
        C->n_kept = 0;
        C->n_kept_nms = 0;
        if (overflow) C->any_overflow = 1;
        C->sum_cand += (long long)n_cand;
    }
}

__global__ void __launch_bounds__(1024)
k2_scan(const int* __restrict__ seg_cnt, int nseg, int* __restrict__ seg_off, int* __restrict__ work_list,
        int* __restrict__ kept_cnt, const unsigned long long* __restrict__ counter, long long cap, PostCtrl* __restrict__ C) {
    pdl_chain_sync();
    int total;
    bool overflow;
    seg_scan_block(seg_cnt, nseg, seg_off, work_list, kept_cnt, counter, cap, C, nullptr, true, &total, &overflow);
}

// scan + bin in one launch (segment tables that fit shared memory): EVERY block scans the counts into its own
// shared-memory table (a few microseconds of redundant L2 reads), block 0 publishes the global results for the kernels
// that follow, then the blocks scatter their share of the keys and boxes.  Saves a launch and the single-CTA critical path.
__global__ void __launch_bounds__(1024)
k2_scan_bin(const int* __restrict__ seg_cnt, int nseg, int* __restrict__ seg_off, int* __restrict__ work_list,
            int* __restrict__ kept_cnt, const unsigned long long* __restrict__ counter, long long cap, PostCtrl* __restrict__ C,
            const uint64_t* __restrict__ keys, const uint32_t* __restrict__ slot, const float4* __restrict__ cbox, int seg_shift,
            uint64_t* __restrict__ bkeys, float4* __restrict__ bbox) {
    pdl_chain_sync();
    extern __shared__ int s_table[];
    int K;
    bool overflow;
    seg_scan_block(seg_cnt, nseg, seg_off, work_list, kept_cnt, counter, cap, C, s_table, blockIdx.x == 0, &K, &overflow);
    __syncthreads();
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < K; p += gridDim.x * blockDim.x) {
        const uint64_t key = keys[p];
        const int dst = s_table[(int)(key >> seg_shift)] + (int)slot[p];
        bkeys[dst] = key;
        bbox[dst] = cbox[p];
    }
}

__global__ void __launch_bounds__(256)
k2_bin(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ slot, const float4* __restrict__ cbox,
       const int* __restrict__ seg_off, int seg_shift, const PostCtrl* __restrict__ C, uint64_t* __restrict__ bkeys,
       float4* __restrict__ bbox) {
    pdl_chain_sync();
    const int K = C->K;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < K; p += gridDim.x * blockDim.x) {
        const uint64_t key = keys[p];
        const int dst = seg_off[(int)(key >> seg_shift)] + (int)slot[p];
        bkeys[dst] = key;
        bbox[dst] = cbox[p];
    }
}

// ------------------------------------------------------------------------------------------ warp NMS by selection
struct SegArgs {
    CandSource src;
    KeyLayout kl;
    float thr;
    int nseg;
    int tiled;
    const TileGeo* geo;
    StitchArgs S;
    const float4* bbox;       // decoded boxes of the binned candidates (same positions as bkeys until a CTA sorts its segment)
};

template <int PL>
__device__ __forceinline__ void warp_select_nms(const SegArgs& A, const uint64_t* __restrict__ bkeys, int s0, int m, int seg, int lane,
                                                float4* __restrict__ rbox, uint64_t* __restrict__ rkey, int* __restrict__ kept_cnt,
                                                int* __restrict__ n_nms) {
    // sub-key = (score field | row): everything below the segment bits, < 2^57.  Split into two 32-bit halves so that
    // the warp minimum is two hardware reductions (REDUX) instead of a 5-step 64-bit shuffle tree.
    const uint64_t sub_mask = (1ull << A.kl.seg_shift) - 1ull;
    uint64_t sub[PL];
    float4 b[PL];
    float a[PL];
    unsigned alive = 0, odd = 0;                                 // odd: a non-finite coordinate / area -> exact IoU path
    const bool thr_plain = A.thr > 0.f && A.thr < 1e30f;
#pragma unroll
    for (int k = 0; k < PL; ++k) {
        const int j = k * 32 + lane;
        sub[k] = ~0ull;
        b[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        a[k] = 0.f;
        if (j < m) {
            sub[k] = bkeys[s0 + j] & sub_mask;
            b[k] = A.bbox[s0 + j];
            a[k] = box_area_exact(b[k]);
            alive |= 1u << k;
            if (!thr_plain || !box_is_plain(b[k], a[k])) odd |= 1u << k;
        }
    }
    TileGeo g{};
    if (A.tiled) g = A.geo[A.src.seg_image(seg)];
    int t = 0, tn = 0;
    while (true) {
        uint64_t best = ~0ull;
        int bk = 0;
#pragma unroll
        for (int k = 0; k < PL; ++k)
            if (((alive >> k) & 1u) && sub[k] < best) { best = sub[k]; bk = k; }
        const uint32_t hi = (uint32_t)(best >> 32), lo = (uint32_t)best;
        const uint32_t mh = __reduce_min_sync(0xffffffffu, hi);
        if (mh == 0xffffffffu) break;                            // nothing alive anywhere
        const uint32_t ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
        const bool mine = (hi == mh) && (lo == ml);              // sub-keys are unique inside a segment
        const int owner = __ffs(__ballot_sync(0xffffffffu, mine)) - 1;
        float4 sel = make_float4(0.f, 0.f, 0.f, 0.f);
        float sa = 0.f;
        int sodd = 0;
#pragma unroll
        for (int k = 0; k < PL; ++k)
            if (k == bk) { sel = b[k]; sa = a[k]; sodd = (odd >> k) & 1u; }
        if (mine) alive &= ~(1u << bk);
        float4 bi;
        bi.x = __shfl_sync(0xffffffffu, sel.x, owner);
        bi.y = __shfl_sync(0xffffffffu, sel.y, owner);
        bi.z = __shfl_sync(0xffffffffu, sel.z, owner);
        bi.w = __shfl_sync(0xffffffffu, sel.w, owner);
        const float ai = __shfl_sync(0xffffffffu, sa, owner);
        const bool pick_odd = __shfl_sync(0xffffffffu, sodd, owner) != 0;
        bool emit = true;
        if (A.tiled) { int4 ib; emit = stitch_box(bi, g, A.S, &ib); }
        if (emit) {
            if (lane == 0) { rbox[s0 + t] = bi; rkey[s0 + t] = ((uint64_t)seg << A.kl.seg_shift) | ((uint64_t)mh << 32) | ml; }
            ++t;
        }
        ++tn;
#pragma unroll
        for (int k = 0; k < PL; ++k) {
            if (!((alive >> k) & 1u)) continue;
            const bool kill = (pick_odd || ((odd >> k) & 1u)) ? suppresses_exact(bi, ai, b[k], a[k], A.thr)
                                                              : suppresses_plain(bi, ai, b[k], a[k], A.thr);
            if (kill) alive &= ~(1u << k);
        }
    }
    if (lane == 0) kept_cnt[seg] = t;
    *n_nms += tn;
}

// Warp kernel: the warps walk the warp-resolved part of the size-ordered work list round-robin (no queue: ~5000
// same-address atomics of a dynamic cursor cost more than the imbalance they remove).
#ifndef K2_WARP_MINB
#define K2_WARP_MINB 4
#endif
__global__ void __launch_bounds__(256, K2_WARP_MINB)
k2_nms_warp(const SegArgs A, const uint64_t* __restrict__ bkeys, const int* __restrict__ seg_off, const int* __restrict__ work_list,
            float4* __restrict__ rbox, uint64_t* __restrict__ rkey, int* __restrict__ kept_cnt, PostCtrl* __restrict__ C) {
    pdl_chain_sync();
    __shared__ int s_nms;
    const int lane = threadIdx.x & 31;
    const int warp_g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    const int first = C->n_mid, last = C->n_small;
    if (threadIdx.x == 0) s_nms = 0;
    __syncthreads();
    int n_nms = 0;
    for (int i = first + warp_g; i < last; i += n_warps) {
        const int seg = work_list[i];
        const int s0 = seg_off[seg];
        const int m = seg_off[seg + 1] - s0;
        if (m <= 32) warp_select_nms<1>(A, bkeys, s0, m, seg, lane, rbox, rkey, kept_cnt, &n_nms);
        else if (m <= 64) warp_select_nms<2>(A, bkeys, s0, m, seg, lane, rbox, rkey, kept_cnt, &n_nms);
        else warp_select_nms<4>(A, bkeys, s0, m, seg, lane, rbox, rkey, kept_cnt, &n_nms);
    }
    if (lane == 0 && n_nms) atomicAdd(&s_nms, n_nms);
    __syncthreads();
    if (threadIdx.x == 0 && s_nms) atomicAdd(&C->n_kept_nms, s_nms);
}

// ------------------------------------------------------------------------------------------ CTA path: sort + chunked NMS
// Bitonic sort (ascending) of m keys in shared memory, "flip" formulation: every compare-exchange puts the smaller
// key at the lower index, so the padding up to the next power of two can stay virtual - a pair whose upper index is
// >= m would compare against +inf and never swaps, and is simply skipped.
__device__ __forceinline__ void cta_bitonic_sort(uint64_t* __restrict__ keys, int m) {
    int n_pad = 1;
    while (n_pad < m) n_pad <<= 1;
    const int half = n_pad >> 1;
    for (int k = 2; k <= n_pad; k <<= 1) {
        const int hk = k >> 1;
        for (int idx = threadIdx.x; idx < half; idx += blockDim.x) {
            const int blk = idx / hk, off = idx - blk * hk;
            const int i = blk * k + off, j = blk * k + k - 1 - off;
            if (j < m) {
                const uint64_t x = keys[i], y = keys[j];
                if (y < x) { keys[i] = y; keys[j] = x; }
            }
        }
        __syncthreads();
        for (int s = k >> 2; s >= 1; s >>= 1) {
            for (int idx = threadIdx.x; idx < half; idx += blockDim.x) {
                const int grp = idx / s, off = idx - grp * s;
                const int i = grp * 2 * s + off, j = i + s;
                if (j < m) {
                    const uint64_t x = keys[i], y = keys[j];
                    if (y < x) { keys[i] = y; keys[j] = x; }
                }
            }
            __syncthreads();
        }
    }
}

// CTA kernel: one CTA per segment, pulled through a device-side cursor over a range of the size-ordered work list
// (a few hundred items at most, largest first).  Launched twice: the rare segments above 4096 boxes with a key buffer
// sized for the largest possible segment of the source, everything else with the shared memory the NMS phase needs
// anyway, two CTAs per SM.
__global__ void __launch_bounds__(NMS_THREADS, 2)
k2_nms_cta(const SegArgs A, uint64_t* __restrict__ bkeys, const int* __restrict__ seg_off, const int* __restrict__ work_list,
           int* __restrict__ cursor, const int* __restrict__ range_end,
           float4* __restrict__ sbox, float* __restrict__ sarea, uint8_t* __restrict__ supp, uint8_t* __restrict__ keepf,
           float4* __restrict__ rbox, uint64_t* __restrict__ rkey, int* __restrict__ kept_cnt, PostCtrl* __restrict__ C) {
    pdl_chain_sync();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* s_keys = reinterpret_cast<uint64_t*>(smem_raw);       // sort phase
    ChunkSmem& S = *reinterpret_cast<ChunkSmem*>(smem_raw);          // NMS phase (the keys are in global memory again by then)
    __shared__ int s_idx[NMS_T];
    __shared__ int s_wcount[NMS_THREADS / 32];
    __shared__ int s_n, s_item, s_run;
    __shared__ long long s_pos;
    const int n_cta_items = *range_end;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    while (true) {
        if (threadIdx.x == 0) s_item = atomicAdd(cursor, 1);
        __syncthreads();
        const int item = s_item;
        if (item >= n_cta_items) break;
        const int seg = work_list[item];
        const int s0 = seg_off[seg];
        const int m = seg_off[seg + 1] - s0;
        const int img = A.src.seg_image(seg);
        // 1. sort the segment's keys
        for (int j = threadIdx.x; j < m; j += NMS_THREADS) s_keys[j] = bkeys[s0 + j];
        __syncthreads();
        cta_bitonic_sort(s_keys, m);
        // 2. sorted keys back, boxes / areas gathered in sorted order
        for (int j = threadIdx.x; j < m; j += NMS_THREADS) {
            const uint64_t key = s_keys[j];
            bkeys[s0 + j] = key;
            const float4 bx = cand_box(A.src, img, (int64_t)(key & A.kl.row_mask));
            sbox[s0 + j] = bx;
            sarea[s0 + j] = box_area_exact(bx);
            supp[s0 + j] = 0;
            keepf[s0 + j] = 0;
        }
        if (threadIdx.x == 0) s_pos = 0;
        __syncthreads();
        // 3. greedy NMS: steps of <= NMS_T still-alive boxes behind a cursor (same routine as k_nms_segments)
        while (true) {
            const int ct = gather_alive(supp, s0, (long long)m, &s_pos, &s_n, s_idx, s_wcount);
            if (ct == 0) break;
            chunk_resolve(S, sbox, sarea, supp, s0, ct, A.thr, s_idx);
            const int nk = S.nk;
            for (int t = threadIdx.x; t < nk; t += NMS_THREADS) keepf[s0 + s_idx[S.klist[t]]] = 1;
            const long long first = s_pos;
            for (long long j = first + threadIdx.x; j < m; j += NMS_THREADS) {
                if (supp[s0 + j]) continue;
                const float4 bj = sbox[s0 + j];
                const float aj = sarea[s0 + j];
                for (int t = 0; t < nk; ++t) {
                    const int k = S.klist[t];
                    if (suppresses_exact(S.box[k], S.area[k], bj, aj, A.thr)) { supp[s0 + j] = 1; break; }
                }
            }
            __syncthreads();
        }
        // 4. ordered compaction of the kept boxes (tiled: only the ones this tile owns) to the front of the segment's range
        TileGeo g{};
        if (A.tiled) g = A.geo[img];
        if (threadIdx.x == 0) s_run = 0;
        __syncthreads();
        int n_nms = 0;
        for (int base = 0; base < m; base += NMS_THREADS) {
            const int j = base + threadIdx.x;
            bool kept = j < m && keepf[s0 + j];
            float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
            n_nms += kept ? 1 : 0;
            if (kept) {
                bx = sbox[s0 + j];
                if (A.tiled) { int4 ib; kept = stitch_box(bx, g, A.S, &ib); }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, kept);
            if (lane == 0) s_wcount[wid] = __popc(bal);
            __syncthreads();
            int before = 0, total = 0;
            for (int w = 0; w < NMS_THREADS / 32; ++w) { const int c = s_wcount[w]; if (w < wid) before += c; total += c; }
            const int run = s_run;
            if (kept) {
                const int q = s0 + run + before + __popc(bal & ((1u << lane) - 1u));
                rbox[q] = bx;
                rkey[q] = bkeys[s0 + j];
            }
            __syncthreads();
            if (threadIdx.x == 0) s_run = run + total;
            __syncthreads();
        }
        for (int o = 16; o; o >>= 1) n_nms += __shfl_xor_sync(0xffffffffu, n_nms, o);
        if (lane == 0 && n_nms) atomicAdd(&C->n_kept_nms, n_nms);
        if (threadIdx.x == 0) kept_cnt[seg] = s_run;
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------ output
__global__ void __launch_bounds__(1024)
k2_out_scan(const int* __restrict__ kept_cnt, int nseg, int* __restrict__ out_off, PostCtrl* __restrict__ C, int tiled) {
    pdl_chain_sync();
    __shared__ int s_warp[33];
    const int tid = threadIdx.x;
    const int per = (nseg + 1023) / 1024;
    const int b0 = tid * per;
    int sum = 0;
    for (int i = 0; i < per; ++i) if (b0 + i < nseg) sum += kept_cnt[b0 + i];
    int total = 0;
    int run = block_excl_scan_1024(sum, s_warp, &total);
    for (int i = 0; i < per; ++i)
        if (b0 + i < nseg) { out_off[b0 + i] = run; run += kept_cnt[b0 + i]; }
    if (tid == 0) {
        out_off[nseg] = total;
        C->n_kept = total;
        C->sum_kept_nms += C->n_kept_nms;
        if (tiled) { C->emit_base = C->acc_rows; C->acc_rows += total; }
    }
}

struct EmitPlain {
    float4* box; float* score; int32_t* label; int32_t* img; int32_t* src;
};

__global__ void __launch_bounds__(256)
k2_emit(const SegArgs A, const int* __restrict__ seg_off, const int* __restrict__ kept_cnt, const int* __restrict__ out_off,
        const float4* __restrict__ rbox, const uint64_t* __restrict__ rkey, const PostCtrl* __restrict__ C, EmitPlain P,
        double* __restrict__ preds, long long cap_rows) {
    pdl_chain_sync();
    const int lane = threadIdx.x & 31;
    const int warp_g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    const long long base = A.tiled ? C->emit_base : 0;
    for (int seg = warp_g; seg < A.nseg; seg += n_warps) {
        const int n = kept_cnt[seg];
        if (n <= 0) continue;
        const int s0 = seg_off[seg], q0 = out_off[seg];
        const int img = A.src.seg_image(seg), label = A.src.seg_label(seg);
        TileGeo g{};
        if (A.tiled) g = A.geo[img];
        for (int t = lane; t < n; t += 32) {
            const float4 bx = rbox[s0 + t];
            const uint64_t key = rkey[s0 + t];
            const float score = key_score(A.kl, key);
            if (A.tiled) {
                const long long q = base + q0 + t;
                if (q < cap_rows) {
                    int4 ib;
                    stitch_box(bx, g, A.S, &ib);
                    double* o = preds + q * 6;
                    o[0] = ib.x; o[1] = ib.y; o[2] = ib.z; o[3] = ib.w;
                    o[4] = (double)score;
                    o[5] = (double)label;
                }
            } else {
                const int q = q0 + t;
                P.box[q] = bx;
                P.score[q] = score;
                P.label[q] = label;
                P.img[q] = img;
                P.src[q] = (int32_t)(key & A.kl.row_mask);
            }
        }
    }
}

// k2_out_scan + k2_emit in one launch: a block owns 8 consecutive segments (one per warp) and computes the output offset of
// its first segment itself (a block-wide sum over the kept counts before it - a few loads per thread from L2); the block
// that finishes last publishes the totals (and, on the tiled path, advances the row counter that every block has read
// before).  Used when the segment table is small enough for the redundant sums to be cheap.
__global__ void __launch_bounds__(256)
k2_emit_fused(const SegArgs A, const int* __restrict__ seg_off, const int* __restrict__ kept_cnt, const float4* __restrict__ rbox,
              const uint64_t* __restrict__ rkey, PostCtrl* __restrict__ C, EmitPlain P, double* __restrict__ preds, long long cap_rows) {
    pdl_chain_sync();
    __shared__ int s_red[8];
    __shared__ int s_cnt[8];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int first = blockIdx.x * 8;
    const long long base = A.tiled ? C->acc_rows : 0;           // read before this block takes its ticket
    int part = 0;
    for (int s = threadIdx.x; s < first; s += 256) part += kept_cnt[s];
    for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) s_red[wid] = part;
    const int seg = first + wid;
    const int n = seg < A.nseg ? kept_cnt[seg] : 0;
    if (lane == 0) s_cnt[wid] = n;
    __syncthreads();
    int q0 = 0;
    for (int w = 0; w < 8; ++w) q0 += s_red[w];
    for (int w = 0; w < wid; ++w) q0 += s_cnt[w];
    if (n > 0) {
        const int s0 = seg_off[seg];
        const int img = A.src.seg_image(seg), label = A.src.seg_label(seg);
        TileGeo g{};
        if (A.tiled) g = A.geo[img];
        for (int t = lane; t < n; t += 32) {
            const float4 bx = rbox[s0 + t];
            const uint64_t key = rkey[s0 + t];
            const float score = key_score(A.kl, key);
            if (A.tiled) {
                const long long q = base + q0 + t;
                if (q < cap_rows) {
                    int4 ib;
                    stitch_box(bx, g, A.S, &ib);
                    double* o = preds + q * 6;
                    o[0] = ib.x; o[1] = ib.y; o[2] = ib.z; o[3] = ib.w;
                    o[4] = (double)score;
                    o[5] = (double)label;
                }
            } else {
                const int q = q0 + t;
                P.box[q] = bx;
                P.score[q] = score;
                P.label[q] = label;
                P.img[q] = img;
                P.src[q] = (int32_t)(key & A.kl.row_mask);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(&C->emit_ticket, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (s_last) {                                                // every other block has read acc_rows and written its rows
        int tot = 0;
        for (int s = threadIdx.x; s < A.nseg; s += 256) tot += kept_cnt[s];
        for (int o = 16; o; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        if (lane == 0) s_red[wid] = tot;
        __syncthreads();
        if (threadIdx.x == 0) {
            int total = 0;
            for (int w = 0; w < 8; ++w) total += s_red[w];
            C->n_kept = total;
            C->sum_kept_nms += C->n_kept_nms;
            if (A.tiled) { C->emit_base = C->acc_rows; C->acc_rows += total; }
            C->emit_ticket = 0;
        }
    }
}

// ------------------------------------------------------------------------------------------ host
bool PostProc::segmented_ok(const CandSource& src) { return src.rows_per_image <= SEG_MID_MAX; }

static SegArgs seg_args(const CandSource& src, const KeyLayout& kl, float thr, const StitchCtx* st) {
    SegArgs A;
    A.src = src; A.kl = kl; A.thr = thr; A.nseg = (int)src.num_segments();
    A.tiled = st ? 1 : 0;
    A.geo = st ? st->geo : nullptr;
    A.S = st ? st->S : StitchArgs{};
    A.bbox = nullptr;
    return A;
}

static void ensure_ctrl(PostProc* P) {
    if (!P->ctrl.p) {
        P->ctrl.reserve(sizeof(PostCtrl));
        Y3_CUDA(cudaMemsetAsync(P->ctrl.p, 0, sizeof(PostCtrl), P->ctx->stream));
    }
}

void PostProc::segmented_front(const CandSource& src, const KeyLayout& kl, int64_t cap) {
    cudaStream_t st = ctx->stream;
    const int nseg = (int)src.num_segments();
    Y3_CHECK(cap < (1ll << 31), Y3_ERR_UNSUPPORTED, "candidate capacity %lld too large", (long long)cap);
    ensure_ctrl(this);
    seg_cnt.reserve((size_t)(nseg + 1) * 4); seg_off32.reserve((size_t)(nseg + 1) * 4);
    mid_list.reserve((size_t)nseg * 4); kept_cnt.reserve((size_t)nseg * 4);
    out_off.reserve((size_t)(nseg + 1) * 4);
    bkeys.reserve((size_t)cap * 8); bbox.reserve((size_t)cap * 16);
    Y3_CUDA(cudaMemsetAsync(seg_cnt.p, 0, (size_t)(nseg + 1) * 4, st));
    {
        Phase p(ctx, &ctx->timings.ms_decode, "y3:decode_threshold_compact");                  // decode + threshold + compaction kernel alone
        launch_candidates(src, kl, cap, true);
        p.stop();
    }
    static const bool fuse = getenv("Y3_NO_POST_FUSE") == nullptr;
    if (fuse && nseg <= 12000) {
        launch_chained(ctx, k2_scan_bin, ctx->sm_count, 1024, (size_t)(nseg + 1) * 4, st, seg_cnt.as<int>(), nseg, seg_off32.as<int>(),
                       mid_list.as<int>(), kept_cnt.as<int>(), counters.as<unsigned long long>(), (long long)cap, ctrl.as<PostCtrl>(),
                       keys[0].as<uint64_t>(), slot.as<uint32_t>(), cbox.as<float4>(), kl.seg_shift, bkeys.as<uint64_t>(), bbox.as<float4>());
    } else {
        launch_chained(ctx, k2_scan, 1, 1024, 0, st, seg_cnt.as<int>(), nseg, seg_off32.as<int>(), mid_list.as<int>(), kept_cnt.as<int>(),
                       counters.as<unsigned long long>(), (long long)cap, ctrl.as<PostCtrl>());
        launch_chained(ctx, k2_bin, ctx->sm_count * 4, 256, 0, st, keys[0].as<uint64_t>(), slot.as<uint32_t>(), cbox.as<float4>(), seg_off32.as<int>(),
                       kl.seg_shift, ctrl.as<PostCtrl>(), bkeys.as<uint64_t>(), bbox.as<float4>());
    }
}

void PostProc::segmented_nms(const CandSource& src, const KeyLayout& kl, float iou_thr, const StitchCtx* stc) {
    cudaStream_t st = ctx->stream;
    const int64_t cap = capacity(src);
    SegArgs A = seg_args(src, kl, iou_thr, stc);
    A.bbox = bbox.as<float4>();
    rbox.reserve((size_t)cap * 16); rkey.reserve((size_t)cap * 8);
    sbox.reserve((size_t)cap * 16); sarea.reserve((size_t)cap * 4); supp.reserve((size_t)cap); keepf.reserve((size_t)cap);
    PostCtrl* C = ctrl.as<PostCtrl>();
    // The CTA-resolved and the warp-resolved segments are independent: the CTA kernels run on an auxiliary stream
    // (fork after k2_bin, join before k2_out_scan) so that the two tails overlap.
    const bool cta_possible = src.rows_per_image > SEG_WARP_MAX;
    cudaStream_t sa = st;
    if (cta_possible) {
        if (!aux_stream) {
            Y3_CUDA(cudaStreamCreateWithFlags(&aux_stream, cudaStreamNonBlocking));
            Y3_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
            Y3_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
        }
        sa = aux_stream;
        Y3_CUDA(cudaEventRecord(ev_fork, st));
        Y3_CUDA(cudaStreamWaitEvent(sa, ev_fork, 0));
        static bool attr_set[64] = {};
        if (!attr_set[ctx->device & 63]) {
            Y3_CUDA(cudaFuncSetAttribute(k2_nms_cta, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SEG_MID_MAX * 8)));
            attr_set[ctx->device & 63] = true;
        }
        static_assert((size_t)SEG_CTAA_MAX * 8 <= sizeof(ChunkSmem), "class A keys must fit the NMS-phase buffer");
        if (src.rows_per_image > SEG_CTAA_MAX) {
            const size_t smem = std::max((size_t)std::min<int64_t>(src.rows_per_image, SEG_MID_MAX) * 8, sizeof(ChunkSmem));
            k2_nms_cta<<<std::min(A.nseg, ctx->sm_count), NMS_THREADS, smem, sa>>>(
                A, bkeys.as<uint64_t>(), seg_off32.as<int>(), mid_list.as<int>(), &C->big_next, &C->n_big, sbox.as<float4>(), sarea.as<float>(),
                supp.as<uint8_t>(), keepf.as<uint8_t>(), rbox.as<float4>(), rkey.as<uint64_t>(), kept_cnt.as<int>(), C);
            Y3_LAUNCHED(ctx);
        }
        k2_nms_cta<<<std::min(A.nseg, ctx->sm_count * 2), NMS_THREADS, sizeof(ChunkSmem), sa>>>(
            A, bkeys.as<uint64_t>(), seg_off32.as<int>(), mid_list.as<int>(), &C->mid_next, &C->n_mid, sbox.as<float4>(), sarea.as<float>(),
            supp.as<uint8_t>(), keepf.as<uint8_t>(), rbox.as<float4>(), rkey.as<uint64_t>(), kept_cnt.as<int>(), C);
        Y3_LAUNCHED(ctx);
        Y3_CUDA(cudaEventRecord(ev_join, sa));
    }
    {
        const int blocks = std::min((A.nseg + 7) / 8, ctx->sm_count * K2_WARP_MINB);
        launch_chained(ctx, k2_nms_warp, blocks, 256, 0, st, A, bkeys.as<uint64_t>(), seg_off32.as<int>(), mid_list.as<int>(), rbox.as<float4>(),
                       rkey.as<uint64_t>(), kept_cnt.as<int>(), C);
    }
    if (cta_possible) Y3_CUDA(cudaStreamWaitEvent(st, ev_join, 0));
    static const bool fuse = getenv("Y3_NO_POST_FUSE") == nullptr;
    fused_emit = fuse && A.nseg <= 16384;
    if (fused_emit) {
        if (stc)
            launch_chained(ctx, k2_emit_fused, (A.nseg + 7) / 8, 256, 0, st, A, seg_off32.as<int>(), kept_cnt.as<int>(), rbox.as<float4>(),
                           rkey.as<uint64_t>(), ctrl.as<PostCtrl>(), EmitPlain{}, stc->preds, (long long)stc->cap_rows);
        return;                                                  // (plain: segmented_emit_plain launches it with the output arrays)
    }
    launch_chained(ctx, k2_out_scan, 1, 1024, 0, st, kept_cnt.as<int>(), A.nseg, out_off.as<int>(), ctrl.as<PostCtrl>(), A.tiled);
    if (stc) {
        const int blocks = std::min((A.nseg + 7) / 8, ctx->sm_count * 8);
        launch_chained(ctx, k2_emit, blocks, 256, 0, st, A, seg_off32.as<int>(), kept_cnt.as<int>(), out_off.as<int>(), rbox.as<float4>(),
                       rkey.as<uint64_t>(), ctrl.as<PostCtrl>(), EmitPlain{}, stc->preds, (long long)stc->cap_rows);
    }
}

void PostProc::segmented_emit_plain(const CandSource& src, const KeyLayout& kl) {
    cudaStream_t st = ctx->stream;
    // the emit kernel and the read-back of the control block are enqueued before anything is waited for (outputs sized
    // for the candidate capacity: a kept box is a candidate) - the device never waits for the host inside one run
    const int64_t cap = capacity(src);
    o_box.reserve(cap * 16); o_score.reserve(cap * 4); o_label.reserve(cap * 4); o_img.reserve(cap * 4); o_src.reserve(cap * 4);
    const SegArgs A = seg_args(src, kl, 0.f, nullptr);
    EmitPlain P{o_box.as<float4>(), o_score.as<float>(), o_label.as<int32_t>(), o_img.as<int32_t>(), o_src.as<int32_t>()};
    if (fused_emit) {
        launch_chained(ctx, k2_emit_fused, (A.nseg + 7) / 8, 256, 0, st, A, seg_off32.as<int>(), kept_cnt.as<int>(), rbox.as<float4>(),
                       rkey.as<uint64_t>(), ctrl.as<PostCtrl>(), P, (double*)nullptr, (long long)0);
    } else {
        const int blocks = std::min((A.nseg + 7) / 8, ctx->sm_count * 8);
        launch_chained(ctx, k2_emit, blocks, 256, 0, st, A, seg_off32.as<int>(), kept_cnt.as<int>(), out_off.as<int>(), rbox.as<float4>(),
                       rkey.as<uint64_t>(), ctrl.as<PostCtrl>(), P, (double*)nullptr, (long long)0);
    }
    host_ctrl.reserve(sizeof(PostCtrl));
    Y3_CUDA(cudaMemcpyAsync(host_ctrl.p, ctrl.p, sizeof(PostCtrl), cudaMemcpyDeviceToHost, st));
    pending_cap = cap;
}

// second half of a plain segmented run: waits for the stream and turns the control block into the result
NmsResult PostProc::finish() {
    NmsResult R = pending;
    if (pending_cap < 0) return R;                               // the run already completed (global-sort path / empty input)
    Y3_CUDA(cudaStreamSynchronize(ctx->stream));
    const PostCtrl C = *host_ctrl.as<PostCtrl>();
    const int64_t cap = pending_cap;
    pending_cap = -1;
    Y3_CHECK(!C.overflow, Y3_ERR_NOSPACE, "candidate list overflow: %llu candidates, capacity %lld (raise y3_config.max_candidates)",
             C.n_cand, (long long)cap);
    R.n_cand = (int64_t)C.n_cand;
    R.n_kept = C.n_kept;
    if (C.n_kept == 0) return R;
    R.boxes = o_box.as<float4>(); R.scores = o_score.as<float>(); R.labels = o_label.as<int32_t>();
    R.img = o_img.as<int32_t>(); R.src_row = o_src.as<int32_t>();
    return R;
}

void PostProc::begin_tiled() {
    ctrl.reserve(sizeof(PostCtrl));
    Y3_CUDA(cudaMemsetAsync(ctrl.p, 0, sizeof(PostCtrl), ctx->stream));
}

bool PostProc::run_tiled(const CandSource& src, float iou_thr, const StitchCtx& stc) {
    static const bool no_seg = getenv("Y3_NMS_GLOBAL_SORT") != nullptr;
    if (no_seg || !segmented_ok(src) || src.num_segments() >= (1ll << 24)) return false;
    const int64_t rows = src.rows_per_image * src.n_images;
    if (rows <= 0) return true;
    Y3_CHECK(rows < (1ll << 32), Y3_ERR_UNSUPPORTED, "too many rows (%lld)", (long long)rows);
    const KeyLayout kl = key_layout(src);
    if (kl.seg_shift > 62) return false;
    segmented_front(src, kl, capacity(src));
    segmented_nms(src, kl, iou_thr, &stc);
    return true;
}

PostCtrl PostProc::finish_tiled() {
    host_ctrl.reserve(sizeof(PostCtrl));
    Y3_CUDA(cudaMemcpyAsync(host_ctrl.p, ctrl.p, sizeof(PostCtrl), cudaMemcpyDeviceToHost, ctx->stream));
    Y3_CUDA(cudaStreamSynchronize(ctx->stream));
    return *host_ctrl.as<PostCtrl>();
}

}  // namespace y3
