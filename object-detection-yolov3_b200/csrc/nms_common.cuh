// nms_common.cuh - device helpers shared by the two post-processing pipelines (nms.cu: global radix sort, for very
// large segments; nms_seg.cu: the synchronisation-free segmented pipeline).
#pragma once
#include "postproc.cuh"

namespace y3 {

static constexpr int NMS_T = 512;             // boxes per chunk
static constexpr int NMS_W = NMS_T / 64;      // mask words per row
static constexpr int NMS_THREADS = 1024;    // mask + apply phases scale with threads; the sweep is one thread
static constexpr int64_t BIG_SEGMENT = 32768; // segments above this use resolve/apply launches (whole GPU per segment)
static constexpr int SEG_WARP_MAX = 128;      // segmented pipeline: one warp resolves a segment of up to this many boxes

// ------------------------------------------------------------------------------------------
// orderable score bits: ascending unsigned order == ascending float order
__device__ __forceinline__ uint32_t orderable(float s) {
    const uint32_t u = __float_as_uint(s);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_orderable(uint32_t o) {
    const uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
    return __uint_as_float(u);
}


// ------------------------------------------------------------------------------------------
// One chunk (<= NMS_T boxes starting at `base`) of one segment: build the IoU bitmask in shared
// memory, sweep it serially, return the kept local indices in s_klist / s_nk.
struct ChunkSmem {
    float4 box[NMS_T];
    float area[NMS_T];
    unsigned long long mask[NMS_T * NMS_W];
    unsigned long long dead[NMS_W];
    int klist[NMS_T];
    int nk;
};

// idx (optional, shared memory): the chunk is the gathered boxes sbox[base + idx[j]] (all alive) instead of the
// contiguous run sbox[base + j]
__device__ __forceinline__ void chunk_resolve(ChunkSmem& S, const float4* __restrict__ sbox,
                                              const float* __restrict__ sarea,
                                              const uint8_t* __restrict__ supp, int64_t base, int ct, float thr,
                                              const int* idx = nullptr) {
    const int tid = threadIdx.x;
    // load + dead bits (already suppressed by earlier chunks, or past the end)
    uint32_t* dead32 = reinterpret_cast<uint32_t*>(S.dead);
    for (int j = tid; j < NMS_T; j += NMS_THREADS) {
        bool dead = true;
        if (j < ct) {
            const int64_t e = base + (idx ? idx[j] : j);
            S.box[j] = sbox[e];
            S.area[j] = sarea[e];
            dead = idx ? false : (supp[e] != 0);
        }
        const unsigned b = __ballot_sync(0xffffffffu, dead);
        if ((tid & 31) == 0) dead32[j >> 5] = b;
    }
    __syncthreads();
    // bitmask: lanes walk rows i, all lanes of a warp share the word w => box j is a broadcast read
    const int nw = (ct + 63) >> 6;
    for (int idx = tid; idx < NMS_T * nw; idx += NMS_THREADS) {
        const int w = idx / NMS_T;
        const int i = idx - w * NMS_T;
        if (i >= ct || w < (i >> 6)) continue;
        if ((S.dead[i >> 6] >> (i & 63)) & 1ull) continue;          // row never read by the sweep
        const float4 bi = S.box[i];
        const float ai = S.area[i];
        unsigned long long bits = 0;
        const int j0 = w << 6;
        const int jn = min(64, ct - j0);
        for (int b = 0; b < jn; ++b) {
            const int j = j0 + b;
            if (j > i && suppresses_exact(bi, ai, S.box[j], S.area[j], thr)) bits |= (1ull << b);
        }
        S.mask[i * NMS_W + w] = bits;
    }
    __syncthreads();
    if (tid == 0) {
        unsigned long long dead[NMS_W];
#pragma unroll
        for (int w = 0; w < NMS_W; ++w) dead[w] = S.dead[w];
        int nk = 0;
#pragma unroll
        for (int w0 = 0; w0 < NMS_W; ++w0) {
            unsigned long long cur = ~dead[w0];
            while (cur) {
                const int b = __ffsll((long long)cur) - 1;
                const int i = (w0 << 6) + b;
                S.klist[nk++] = i;
                const unsigned long long* row = S.mask + i * NMS_W;
#pragma unroll
                for (int w = 0; w < NMS_W; ++w)
                    if (w >= w0) dead[w] |= row[w];
                const unsigned long long above = (b == 63) ? 0ull : (~0ull << (b + 1));
                cur = ~dead[w0] & above;
            }
        }
        S.nk = nk;
    }
    __syncthreads();
}

// Gathers (in order) the indices of the next <= NMS_T still-alive boxes of a segment behind *s_pos into s_idx and
// advances *s_pos past the last position examined.  All NMS_THREADS threads call it; returns the count.
__device__ __forceinline__ int gather_alive(const uint8_t* __restrict__ supp, int64_t s0, long long m, long long* s_pos, int* s_n,
                                            int* s_idx, int* s_wcount) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) *s_n = 0;
    __syncthreads();
    while (true) {
        const long long pos = *s_pos;
        const int have = *s_n;
        if (pos >= m || have >= NMS_T) break;
        const long long p = pos + threadIdx.x;
        const bool alive = p < m && supp[s0 + p] == 0;
        const unsigned b = __ballot_sync(0xffffffffu, alive);
        if (lane == 0) s_wcount[wid] = __popc(b);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < NMS_THREADS / 32; ++w) { const int c = s_wcount[w]; if (w < wid) before += c; total += c; }
        const int rank = have + before + __popc(b & ((1u << lane) - 1u));
        if (alive && rank < NMS_T) s_idx[rank] = (int)p;
        // positions consumed: the whole window if everything fitted, else up to the alive box that took the last slot
        __syncthreads();
        if (have + total <= NMS_T) {
            if (threadIdx.x == 0) { *s_n = have + total; *s_pos = min(pos + (long long)NMS_THREADS, m); }
        } else {
            if (alive && rank == NMS_T - 1) { *s_n = NMS_T; *s_pos = p + 1; }
        }
        __syncthreads();
    }
    return *s_n;
}


// box of a candidate: decoded on the fly from the raw heads, or read from the decoded rows
__device__ __forceinline__ float4 cand_box(const CandSource& src, int64_t img, int64_t row) {
    if (src.from_heads) {
        int sc, cell, a;
        const float* ho;
        const float* hp = head_row(src.dec, (int)img, (int)row, &sc, &cell, &a, &ho);
        return decode_box(src.dec, hp, sc, cell, a);
    }
    const float* b = src.box + (img * src.rows_per_image + row) * src.box_stride;
    return make_float4(__ldg(b), __ldg(b + 1), __ldg(b + 2), __ldg(b + 3));
}
__device__ __forceinline__ float key_score(const KeyLayout& kl, uint64_t key) {
    return from_orderable(~(((uint32_t)(key >> kl.row_bits) & kl.score_mask) + kl.score_base));
}

}  // namespace y3
