// conv_stem1.cu - the first two conv_layers of the backbone (model.py:385-386) as ONE kernel for 1-channel
// images:   stem 3x3 s1 (1 -> 32) -> leaky -> BN      then      3x3 s2 (32 -> 64) -> leaky -> BN.
//
// Why: both layers are HBM-bound on the stem's 32-channel full-resolution activation (16.8 MB per 512x512 tile,
// written once and read once = 4.3 GB of the 35.9 GB a 128-tile batch moves, 1.2 of 13 ms).  Here that tensor
// never leaves the SM.  Both convolutions run on the tensor cores:
//
//   stem warps (8)   gather: thread = one stem column, two stem rows per step; the 9 taps (+ two 1.0 columns that
//                    carry the bias, as in stem_tc.cu) go as bf16 rows into four NO-SWIZZLE K-major A tiles
//                    (row parity x column parity, 128 pixels x K = 16 each)
//   stem MMA thread  stem UMMAs (M = 256 per CTA pair, N = 32, K = 16) into TMEM
//   stem warps       drain the stem accumulators: leaky -> scale/shift -> bf16 -> straight into the K-major
//                    SWIZZLE_64B ring the conv2d_1 UMMAs read (stem pixels outside the image are written as zeros:
//                    they are conv2d_1's "SAME" padding, 0 before / 1 after)
//   conv MMA thread  conv2d_1 UMMAs of the tile whose ring rows are complete (weights stationary, taps by
//                    descriptor, as conv_halo.cu)
//   epilogue warps   conv2d_1: bias -> leaky -> BN -> bf16 -> staging -> TMA store
//
//   * a conv2d_1 tile = 128 output pixels of one row; it needs stem rows 2h, 2h+1, 2h+2 and stem columns
//     2*w0 .. 2*w0+256.  A ring slot holds one stem row split by column parity: E[129][32] (even columns) and
//     O[128][32] (odd columns), so that tap dx = 0 / 1 / 2 is E, O, E shifted by one row of the swizzled tile.
//     Going down a column of tiles, step g of a run produces stem rows 2(h0+g)-1 and 2(h0+g) (only the second one
//     for g = 0), i.e. each new tile costs two new stem rows.  The 257th column of a ring row (E row 128) is
//     computed by a warp of its own on the FP32 pipe, one channel per lane.
//   * two CTAs (images 2*ip and 2*ip+1, same column, same rows) form every UMMA (cta_group::2); each holds half
//     of both weight matrices.  Barriers that collect generic-proxy writes of both CTAs live in the leader.
//   * the step loop is a chain of latencies, not of work (a clock64 trace of the first version: 5150 cycles per
//     step, of which the stem warps spent 1650 waiting for their own UMMA, the MMA thread 2600 issuing 22 UMMAs
//     and the 257th column 1050 on the critical warp).  So: the stem warps write the A rows of step g+1 BEFORE
//     they drain step g (taps prefetched one step ahead), the two UMMA streams are issued by two threads, and
//     the 257th column has its own warp.
#include "conv_tc.cuh"
#include "ptx.cuh"
#include <stdlib.h>
#include <algorithm>

namespace y3 {
using namespace ptx;

static constexpr int SC_STEM_WARPS = 8;
static constexpr int SC_EPI_WARPS = 8;                    // two groups of 4: group e drains the tiles with (tile & 1) == e
static constexpr int SC_THREADS = 64 + 32 * SC_EPI_WARPS + 32 * SC_STEM_WARPS + 32;   // + the 257th-column warp
static constexpr int SC_FULL_ARRIVALS = 2 * (SC_STEM_WARPS + 1);
static constexpr int SC_RING_SLOTS = 7;                   // ring slots (stem rows)

struct SCfg {
    static constexpr int CIN = 32, COUT = 64, ROWB = 64;
    static constexpr int E_BYTES = 9216;                  // 129 rows x 64 B, padded to 1024
    static constexpr int O_BYTES = 8192;                  // 128 rows x 64 B
    static constexpr int SLOT = E_BYTES + O_BYTES;
    static constexpr int S = SC_RING_SLOTS;
    static constexpr int WTAP = (COUT / 2) * ROWB;        // this CTA's half of one conv2d_1 tap
    static constexpr int W_BYTES = 9 * WTAP;
    static constexpr int STG_BYTES = 128 * COUT * 2;
    static constexpr int PAR_FLOATS = 3 * COUT + 9 * 32 + 3 * 32;   // conv2d_1 b|s|t, stem weights, stem b|s|t
    static constexpr int SA_TILE = 128 * 32;              // stem A tile: 128 pixels x K = 16 bf16, no swizzle
    static constexpr int SA_BUF = 4 * SA_TILE;            // (row parity, column parity)
    static constexpr int SB_BYTES = 16 * 32;              // this CTA's half of the stem weights [32][16]
    static constexpr int BAR_BYTES = (2 * S + 13) * 8 + 16;
    static constexpr int SMEM = 1024 + W_BYTES + S * SLOT + 2 * STG_BYTES + 2 * SA_BUF + SB_BYTES + PAR_FLOATS * 4 + BAR_BYTES;
    static constexpr uint32_t TMEM_COLS = 512;            // conv2d_1: 2 x 64; stem: 2 buffers x 4 tiles x 32
    static constexpr uint32_t STEM_COL0 = 2 * COUT;
    static constexpr uint32_t SBO = 8 * ROWB;
    static_assert(SLOT % 1024 == 0 && WTAP % 1024 == 0, "operand tiles must stay 1024-B aligned");
    static_assert(SMEM <= 232448 - 1024, "exceeds 227 KB of shared memory");
};

__device__ __forceinline__ uint32_t pack2s(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint64_t make_smem_desc_interleaved(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;      // K-direction core-matrix stride
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;      // M/N-direction 8-row group stride
    d |= (uint64_t)1 << 46;                                  // descriptor version (Blackwell)
    return d;                                                // layout_type 0 = no swizzle
}

// Passed by value (__grid_constant__): the stem parameters travel in the kernel parameters.
#ifdef Y3_STEM_TRACE
#define TR(role, step, pt) do { if (dbgp && blockIdx.x == 0 && (step) >= 64 && (step) < 80) dbgp[((role) * 16 + ((step) - 64)) * 8 + (pt)] = clock64(); } while (0)
#else
#define TR(role, step, pt) do { } while (0)
#endif
struct StemArgs {
    float w[9 * 32];          // [tap][channel]
    float bias[32], scale[32], shift[32];
    const float* in;          // [n_img][H][W] fp32 (one channel)
    int H, W;
    long long* dbg;           // Y3_STEM_TRACE builds only
};

// the 4 x 3 image window of two vertically adjacent stem pixels (rows y0, y0+1) in column c (zeros outside the image)
__device__ __forceinline__ void stem_taps(const float* __restrict__ imgp, bool img_ok, int H, int W, int y0, int c, float (&tp)[4][3]) {
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
        const int yy = y0 - 1 + rr;
        const bool yok = img_ok && yy >= 0 && yy < H;
        const float* pr = imgp + (long long)yy * W + c;
        tp[rr][0] = (yok && c - 1 >= 0 && c - 1 < W) ? __ldg(pr - 1) : 0.f;
        tp[rr][1] = (yok && c < W) ? __ldg(pr) : 0.f;
        tp[rr][2] = (yok && c + 1 < W) ? __ldg(pr + 1) : 0.f;
    }
}

// one channel (= lane) of the stem pixels (y0, c) and (y0+1, c) on the FP32 pipe: the 257th pixel of a ring row
__device__ __forceinline__ void stem_pixel_by_lanes(const float* __restrict__ imgp, bool img_ok, int H, int W, int y0, int c,
                                                    const float* __restrict__ s_w, const float* __restrict__ s_sp, int lane,
                                                    unsigned char* const (&rowp)[2], bool skip_first) {
    float acc[2] = {0.f, 0.f};
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
        const int yy = y0 - 1 + rr;
        const bool yok = img_ok && yy >= 0 && yy < H;
        const float* pr = imgp + (long long)yy * W + c;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            const int xx = c - 1 + dx;
            const float x = (yok && xx >= 0 && xx < W) ? __ldg(pr - 1 + dx) : 0.f;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int dy = rr - q;
                if (dy >= 0 && dy < 3) acc[q] = fmaf(s_w[(dy * 3 + dx) * 32 + lane], x, acc[q]);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        if (q == 0 && skip_first) continue;
        float z = acc[q] + s_sp[lane];
        z = fmaxf(z, 0.2f * z) * s_sp[32 + lane] + s_sp[64 + lane];
        const bool inside = (y0 + q >= 0) && (y0 + q < H) && (c < W);
        const __nv_bfloat16 v = __float2bfloat16_rn(inside ? z : 0.f);
        // row 128 of the E array: swizzle term (128 >> 1) & 3 = 0
        *reinterpret_cast<__nv_bfloat16*>(rowp[q] + ((lane >> 3) << 4) + (lane & 7) * 2) = v;
    }
}

// Walks the stem steps of one CTA pair in order.  The pair's conv2d_1 tiles [t_begin, t_end) (index = column * Ho + row)
// fall into runs down one column; a run of len tiles has len + 1 steps: step g produces the stem rows 2(h0+g)-1 and
// 2(h0+g) (g = 0: only the second) into the ring rows k (and k + 1).
struct Steps {
    int t_end, Ho, tiles_x;
    int t, h0, len, g;         // run: first tile, its row, tiles; position inside the run
    int ip, w0;                // run: image pair, first conv2d_1 output column
    int gs, k;                 // global step counter; ring row index of the step's first produced row
    int slot, par;             // k % S and (k / S) & 1, kept incrementally (S = ring slots)
    __device__ __forceinline__ void start_run() {
        if (t < t_end) {
            const int col = t / Ho;
            h0 = t - col * Ho;
            len = min(t_end - t, Ho - h0);
            ip = col / tiles_x;
            w0 = (col - ip * tiles_x) * 128;
            g = 0;
        }
    }
    __device__ __forceinline__ void init(int t_begin, int t_end_, int Ho_, int tiles_x_) {
        t_end = t_end_; Ho = Ho_; tiles_x = tiles_x_; t = t_begin; gs = 0; k = 0; slot = 0; par = 0;
        h0 = len = g = ip = w0 = 0;
        start_run();
    }
    __device__ __forceinline__ bool done() const { return t >= t_end; }
    __device__ __forceinline__ void next() {
        const int n = g == 0 ? 1 : 2;
        k += n;
        slot += n;
        if (slot >= SC_RING_SLOTS) { slot -= SC_RING_SLOTS; par ^= 1; }
        ++gs;
        if (++g > len) { t += len; start_run(); }
    }
    __device__ __forceinline__ int first() const { return g == 0; }
    __device__ __forceinline__ int y0() const { return 2 * (h0 + g) - 1; }      // stem rows y0 (unused when first) and y0 + 1
};

// what the stem warps remember about the step whose accumulators they still have to drain
struct StepInfo {
    int valid, buf, y0, w0, first;
    int slot, par;                         // ring slot / use parity of the step's first produced row
    uint32_t use;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(SC_THREADS, 1)
k_stem_conv1(const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_out, const ConvArgs P,
             const __grid_constant__ StemArgs T) {
#if defined(__CUDA_ARCH_FEAT_SM100_ALL) || defined(__CUDA_ARCH_FEAT_SM101_ALL)
    using C = SCfg;
    long long* const dbgp = T.dbg; (void)dbgp;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = align_smem_1024(smem_raw);
    unsigned char* w_base = smem;
    unsigned char* ring = smem + C::W_BYTES;
    unsigned char* stg_base = ring + C::S * C::SLOT;
    unsigned char* sa_base = stg_base + 2 * C::STG_BYTES;                     // stem A tiles, 2 buffers x 4 tiles
    unsigned char* sb_base = sa_base + 2 * C::SA_BUF;                         // stem weights (this CTA's 16 channels)
    float* s_par = reinterpret_cast<float*>(sb_base + C::SB_BYTES);           // conv2d_1 bias | scale | shift
    float* s_w = s_par + 3 * C::COUT;                                         // stem weights [9][32] fp32 (257th pixel)
    float* s_sp = s_w + 9 * 32;                                               // stem bias | scale | shift
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_par + C::PAR_FLOATS);
    uint64_t* full = bars;                       // leader only: the stem warps and the 257th-column warp of both CTAs arrive per ring row
    uint64_t* empty = bars + C::S;               // per CTA, multicast commit
    uint64_t* tmem_full = bars + 2 * C::S;       // per CTA, multicast commit
    uint64_t* tmem_empty = tmem_full + 2;        // leader only
    uint64_t* w_full = tmem_full + 4;            // leader only: both halves of the conv2d_1 weights
    uint64_t* sa_full = tmem_full + 5;           // [2] leader only: stem A tiles written (16 warps)
    uint64_t* sa_empty = tmem_full + 7;          // [2] per CTA, multicast commit: stem A tiles consumed
    uint64_t* sacc_full = tmem_full + 9;         // [2] per CTA, multicast commit: stem accumulators ready
    uint64_t* sacc_empty = tmem_full + 11;       // [2] leader only: stem accumulators drained (16 warps)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 13);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1;
    const int n_pairs = gridDim.x >> 1;

    if (warp == 0 && lane == 0) { prefetch_tmap(&map_b); prefetch_tmap(&map_out); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < C::S; ++s) { mbar_init(&full[s], SC_FULL_ARRIVALS); mbar_init(&empty[s], 1); }
        for (int p = 0; p < 2; ++p) {
            mbar_init(&tmem_full[p], 1); mbar_init(&tmem_empty[p], 8);
            mbar_init(&sa_full[p], 16); mbar_init(&sa_empty[p], 1);
            mbar_init(&sacc_full[p], 1); mbar_init(&sacc_empty[p], 16);
        }
        mbar_init(w_full, 1);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc2<C::TMEM_COLS>(tmem_slot);
    for (int i = threadIdx.x; i < C::COUT; i += SC_THREADS) {
        s_par[i] = P.bias[i]; s_par[C::COUT + i] = P.scale[i]; s_par[2 * C::COUT + i] = P.shift[i];
    }
    for (int i = threadIdx.x; i < 9 * 32; i += SC_THREADS) s_w[i] = T.w[i];
    for (int i = threadIdx.x; i < 32; i += SC_THREADS) { s_sp[i] = T.bias[i]; s_sp[32 + i] = T.scale[i]; s_sp[64 + i] = T.shift[i]; }
    if (threadIdx.x < 32) {
        // stem B operand: rows = this CTA's 16 output channels, K = 16: taps 0..8, bf16(bias), bf16(bias - bf16(bias)), 0...
        // interleaved layout: byte = (n / 8) * 256 + (k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2
        const int nl = threadIdx.x >> 1, kc = threadIdx.x & 1;
        const int n = (int)rank * 16 + nl;
        const float bn = T.bias[n];
        const float bhi = __bfloat162float(__float2bfloat16_rn(bn));
        uint32_t pk[4];
#pragma unroll
        for (int jx = 0; jx < 4; ++jx) {
            float ab[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int k = kc * 8 + 2 * jx + e;
                ab[e] = k < 9 ? T.w[k * 32 + n] : (k == 9 ? bhi : (k == 10 ? bn - bhi : 0.f));
            }
            pk[jx] = pack2s(ab[0], ab[1]);
        }
        *reinterpret_cast<uint4*>(sb_base + (nl >> 3) * 256 + kc * 128 + (nl & 7) * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // pair-tiles: (image pair ip, segment seg, output row h) flattened with h fastest; this pair's range
    const int Ho = P.Ho;
    const int img_pairs = (P.n_img + 1) >> 1;
    const long long total = (long long)img_pairs * P.tiles_x * Ho;
    const int t_begin = (int)(total * pair / n_pairs);
    const int t_end = (int)(total * (pair + 1) / n_pairs);

    if (warp == 0) {
        // ------------------------------------------------------------ conv2d_1 weights (stationary), then the stem UMMAs
        if (lane == 0) {
            const uint32_t lead_w = mapa_u32(w_full, 0);
            if (rank == 0) mbar_expect_tx(w_full, (uint32_t)(2 * C::W_BYTES));
            for (int tap = 0; tap < 9; ++tap)
                tma2_load_2d(w_base + tap * C::WTAP, &map_b, lead_w, tap * C::CIN, (int)rank * (C::COUT / 2));
        }
        if (rank == 0 && lane == 0) {
            constexpr uint32_t idesc_stem = make_idesc_bf16(256, 32);
            const uint64_t sb_desc = make_smem_desc_interleaved(smem_u32(sb_base), 128, 256);
            Steps st;
            for (st.init(t_begin, t_end, Ho, P.tiles_x); !st.done(); st.next()) {
                const int buf = st.gs & 1;
                const uint32_t use = (uint32_t)(st.gs >> 1);
                TR(0, st.gs, 0);
                mbar_wait(&sacc_empty[buf], (use & 1u) ^ 1u);
                mbar_wait(&sa_full[buf], use & 1u);
                TR(0, st.gs, 1);
                tc_fence_after();
#pragma unroll
                for (int tile = 0; tile < 4; ++tile) {
                    const uint64_t adesc = make_smem_desc_interleaved(smem_u32(sa_base + buf * C::SA_BUF + tile * C::SA_TILE), 128, 256);
                    umma2_bf16(tmem_base + C::STEM_COL0 + (uint32_t)(buf * 128 + tile * 32), adesc, sb_desc, idesc_stem, 0u);
                }
                umma2_commit_mc(&sacc_full[buf], 3);
                umma2_commit_mc(&sa_empty[buf], 3);
                TR(0, st.gs, 2);
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ conv2d_1 MMA issuer: leader CTA, one thread
        if (rank == 0 && lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(256, C::COUT);
            mbar_wait(w_full, 0);
            tc_fence_after();
            const uint32_t w_addr = smem_u32(w_base);
            const uint32_t ring_addr = smem_u32(ring);
            int it = 0;                 // conv2d_1 tile counter
            int k0 = 0;                 // ring row index of the current run's first row
            for (int t = t_begin; t < t_end;) {
                const int col = t / Ho;
                const int h0 = t - col * Ho;
                const int len = min(t_end - t, Ho - h0);
                for (int r = 0; r < len; ++r, ++it) {
                    const int p = it & 1;
                    const uint32_t use = (uint32_t)(it >> 1);
                    TR(3, it, 0);
                    mbar_wait(&tmem_empty[p], (use & 1u) ^ 1u);
                    TR(3, it, 1);
                    for (int dy = (r == 0 ? 0 : 1); dy < 3; ++dy) {
                        const int kk = k0 + 2 * r + dy;
                        mbar_wait(&full[kk % C::S], (uint32_t)((kk / C::S) & 1));
                    }
                    TR(3, it, 2);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(p * C::COUT);
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy) {
                        const uint32_t slot_addr = ring_addr + (uint32_t)(((k0 + 2 * r + dy) % C::S) * C::SLOT);
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            // dx = 0: even columns; dx = 1: odd columns; dx = 2: even columns shifted by one pixel
                            const uint32_t a_addr = slot_addr + (dx == 1 ? (uint32_t)C::E_BYTES : (dx == 2 ? (uint32_t)C::ROWB : 0u));
                            const uint64_t adesc = make_smem_desc(a_addr, C::SBO, SWZ_64B);
                            const uint64_t bdesc = make_smem_desc(w_addr + (uint32_t)((dy * 3 + dx) * C::WTAP), C::SBO, SWZ_64B);
#pragma unroll
                            for (int kc = 0; kc < C::CIN / 16; ++kc)
                                umma2_bf16(d_tmem, adesc + (uint64_t)(kc * 2), bdesc + (uint64_t)(kc * 2), idesc, (uint32_t)((dy | dx | kc) != 0));
                        }
                    }
                    umma2_commit_mc(&tmem_full[p], 3);
                    umma2_commit_mc(&empty[(k0 + 2 * r) % C::S], 3);
                    umma2_commit_mc(&empty[(k0 + 2 * r + 1) % C::S], 3);
                    if (r == len - 1) umma2_commit_mc(&empty[(k0 + 2 * r + 2) % C::S], 3);
                    TR(3, it, 3);
                }
                k0 += 2 * len + 1;
                t += len;
            }
        }
    } else if (warp == 2 + SC_EPI_WARPS + SC_STEM_WARPS) {
        // ------------------------------------------------------------ 257th stem column of every ring row (E row 128)
        const int H = T.H, W = T.W;
        Steps st;
        for (st.init(t_begin, t_end, Ho, P.tiles_x); !st.done(); st.next()) {
            const int img = 2 * st.ip + (int)rank;
            const bool img_ok = img < P.n_img;
            const float* imgp = T.in + (long long)(img_ok ? img : 0) * H * W;
            const int first = st.first();
            const int slot0 = st.slot, par0 = st.par;
            int slot1 = slot0, par1 = par0;
            if (!first && ++slot1 == C::S) { slot1 = 0; par1 ^= 1; }
            if (!first) mbar_wait(&empty[slot0], (uint32_t)(par0 ^ 1));
            mbar_wait(&empty[slot1], (uint32_t)(par1 ^ 1));
            unsigned char* const re[2] = {ring + slot0 * C::SLOT + 128 * C::ROWB, ring + slot1 * C::SLOT + 128 * C::ROWB};
            stem_pixel_by_lanes(imgp, img_ok, H, W, st.y0(), 2 * st.w0 + 256, s_w, s_sp, lane, re, first != 0);
            fence_proxy_async_smem();              // generic-proxy writes -> visible to the tensor core's reads
            __syncwarp();
            if (lane == 0) {
                if (!first) mbar_arrive_cluster(mapa_u32(&full[slot0], 0));
                mbar_arrive_cluster(mapa_u32(&full[slot1], 0));
            }
        }
    } else if (warp >= 2 + SC_EPI_WARPS) {
        // ------------------------------------------------------------ stem warps
        const int sw = warp - (2 + SC_EPI_WARPS);
        const int st_col = sw * 32 + lane;                      // 0..255 = stem column offset inside the tile's 257 columns
        const int qd = warp & 3;                                // TMEM lane quadrant of this warp
        const int hi = sw >> 2;                                 // stem row of a step (0 / 1) this warp drains
        const int j = st_col >> 1, par = st_col & 1;
        const int H = T.H, W = T.W;
        const bool tr_on = lane == 0 && (sw == 0 || sw == 7);
        const int tr_role = sw == 0 ? 1 : 2; (void)tr_role; (void)tr_on;
        // drain the accumulators of a finished stem step into the ring
        auto drain = [&](const StepInfo& s, int gs_now) {
            mbar_wait(&sacc_full[s.buf], s.use & 1u);
            if (tr_on) TR(tr_role, gs_now, 2);
            tc_fence_after();
            const bool mine = !(s.first && hi == 0);            // step 0 of a run: its first stem row is not needed
            const uint32_t t_addr = tmem_base + ((uint32_t)(qd * 32) << 16) + C::STEM_COL0 + (uint32_t)(s.buf * 128 + hi * 64);
            // ring rows of this step: [first ? nothing : the row in slot0] , the next row (first: slot0 itself)
            const int slot0 = s.slot, par0 = s.par;
            int slot1 = slot0, par1 = par0;
            if (!s.first && ++slot1 == C::S) { slot1 = 0; par1 ^= 1; }
            if (!s.first) mbar_wait(&empty[slot0], (uint32_t)(par0 ^ 1));
            mbar_wait(&empty[slot1], (uint32_t)(par1 ^ 1));
            unsigned char* const slot_mine = ring + (hi ? slot1 : slot0) * C::SLOT;      // the ring row this warp writes
            if (tr_on) TR(tr_role, gs_now, 3);
            if (mine) {
                const int y = s.y0 + hi;
                const int px = qd * 32 + lane;                  // pixel index inside the E / O array
                const int swz = (px >> 1) & 3;
#pragma unroll 1
                for (int pp = 0; pp < 2; ++pp) {                // even / odd columns
                    uint32_t v[32];
                    tmem_ld_32x32(t_addr + (uint32_t)(pp * 32), v);
                    tmem_ld_wait();
                    if (tr_on) TR(tr_role, gs_now, 4 + 2 * pp);
                    const bool inside = (y < H) && (2 * s.w0 + 2 * px + pp < W);
                    unsigned char* rowp = slot_mine + (pp ? C::E_BYTES : 0) + px * C::ROWB;
#pragma unroll
                    for (int c16 = 0; c16 < 4; ++c16) {
                        uint32_t o[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int ch = c16 * 8 + 2 * e;
                            const float2 z = make_float2(__uint_as_float(v[ch]), __uint_as_float(v[ch + 1]));
                            float2 l = mul2_f32(z, make_float2(0.2f, 0.2f));
                            l.x = fmaxf(l.x, z.x); l.y = fmaxf(l.y, z.y);
                            const float2 yv = fma2_f32(l, make_float2(s_sp[32 + ch], s_sp[33 + ch]), make_float2(s_sp[64 + ch], s_sp[65 + ch]));
                            o[e] = inside ? pack2s(yv.x, yv.y) : 0u;
                        }
                        *reinterpret_cast<uint4*>(rowp + ((c16 ^ swz) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
                    }
                    if (tr_on && pp == 0) TR(tr_role, gs_now, 5);
                }
            }
            tc_fence_before();
            fence_proxy_async_smem();              // generic-proxy writes -> visible to the tensor core's reads
            __syncwarp();
            if (lane == 0) {
                mbar_arrive_cluster(mapa_u32(&sacc_empty[s.buf], 0));
                if (!s.first) mbar_arrive_cluster(mapa_u32(&full[slot0], 0));
                mbar_arrive_cluster(mapa_u32(&full[slot1], 0));
            }
            if (tr_on) TR(tr_role, gs_now, 7);
        };
        auto taps_of = [&](const Steps& q, float (&tp)[4][3]) {
            const int img = 2 * q.ip + (int)rank;
            const bool img_ok = img < P.n_img;
            const float* imgp = T.in + (long long)(img_ok ? img : 0) * H * W;
            stem_taps(imgp, img_ok, H, W, q.y0(), 2 * q.w0 + st_col, tp);
        };
        Steps st;
        st.init(t_begin, t_end, Ho, P.tiles_x);
        StepInfo prev;
        prev.valid = 0;
        float tp[4][3];
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) tp[rr][0] = tp[rr][1] = tp[rr][2] = 0.f;
        if (!st.done()) taps_of(st, tp);
        while (!st.done()) {
            const int buf = st.gs & 1;
            const uint32_t use = (uint32_t)(st.gs >> 1);
            StepInfo cur;
            cur.valid = 1; cur.buf = buf; cur.use = use; cur.y0 = st.y0(); cur.w0 = st.w0; cur.slot = st.slot; cur.par = st.par; cur.first = st.first();
            const int gs_now = st.gs;
            st.next();
            if (tr_on) TR(tr_role, gs_now, 0);
            mbar_wait(&sa_empty[buf], (use & 1u) ^ 1u);
            // A rows of the two stem pixels of this column: K = taps 0..8, 1.0, 1.0, zeros (interleaved layout)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                unsigned char* rowp = sa_base + buf * C::SA_BUF + (q * 2 + par) * C::SA_TILE + (j >> 3) * 256 + (j & 7) * 16;
                *reinterpret_cast<uint4*>(rowp) = make_uint4(pack2s(tp[q][0], tp[q][1]), pack2s(tp[q][2], tp[q + 1][0]),
                                                              pack2s(tp[q + 1][1], tp[q + 1][2]), pack2s(tp[q + 2][0], tp[q + 2][1]));
                *reinterpret_cast<uint4*>(rowp + 128) = make_uint4(pack2s(tp[q + 2][2], 1.0f), pack2s(1.0f, 0.f), 0u, 0u);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(&sa_full[buf], 0));
            if (tr_on) TR(tr_role, gs_now, 1);
            // the NEXT step's taps: requested after the fence above (it would wait for them) and in flight while the
            // previous step is drained
            float tn[4][3];
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) tn[rr][0] = tn[rr][1] = tn[rr][2] = 0.f;
            if (!st.done()) taps_of(st, tn);
            if (prev.valid) drain(prev, gs_now);                // the previous step: its UMMAs were issued an iteration ago
            prev = cur;
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) { tp[rr][0] = tn[rr][0]; tp[rr][1] = tn[rr][1]; tp[rr][2] = tn[rr][2]; }
        }
        if (prev.valid) drain(prev, -1);
        // tail: multicast commits from the leader may still arrive on this CTA's barriers
        if (sw == 0 && lane == 0) {
            int k = st.k;
            for (int jx = 0; jx < C::S; ++jx, ++k) mbar_wait(&empty[k % C::S], (uint32_t)(((k / C::S) & 1) ^ 1));
            for (int b = 0; b < 2; ++b) {
                const int gnext = st.gs + (((st.gs & 1) != b) ? 1 : 0);      // next step that would use A buffer b
                mbar_wait(&sa_empty[b], (uint32_t)(((gnext >> 1) & 1) ^ 1));
            }
        }
    } else {
        // ------------------------------------------------------------ conv2d_1 epilogue: warps 2..5 of each CTA
        const int q = warp & 3;                    // TMEM lane quadrant
        const int eg = (warp - 2) >> 2;            // epilogue group = accumulator / staging buffer it owns
        const bool elected = ((warp - 2) & 3) == 0 && lane == 0;
        const int row = q * 32 + lane;
        int it = 0;
        for (int t = t_begin; t < t_end; ++t, ++it) {
            if ((it & 1) != eg) continue;
            const int col = t / Ho;
            const int h = t - col * Ho;
            const int ip = col / P.tiles_x;
            const int w0 = (col - ip * P.tiles_x) * 128;
            const int img = 2 * ip + (int)rank;
            const int p = it & 1;
            const uint32_t use = (uint32_t)(it >> 1);
            if (elected && eg == 0) TR(4, it, 0);
            mbar_wait(&tmem_full[p], use & 1u);
            if (elected && eg == 0) TR(4, it, 1);
            tc_fence_after();
            unsigned char* stg = stg_base + p * C::STG_BYTES;
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(p * C::COUT);
#pragma unroll 1
            for (int g = 0; g < 2; ++g) {            // two 32-column groups of the 64 output channels
                uint32_t v[32];
                tmem_ld_32x32(t_addr + (uint32_t)(g * 32), v);
                tmem_ld_wait();
                const int c0 = g * 32;
                float y[32];
                const float4* pb = reinterpret_cast<const float4*>(s_par + c0);
                const float4* ps = reinterpret_cast<const float4*>(s_par + C::COUT + c0);
                const float4* pt = reinterpret_cast<const float4*>(s_par + 2 * C::COUT + c0);
#pragma unroll
                for (int k4 = 0; k4 < 8; ++k4) {
                    const float4 b4 = pb[k4], s4 = ps[k4], t4 = pt[k4];
                    const float2 za = add2_f32(make_float2(__uint_as_float(v[4 * k4 + 0]), __uint_as_float(v[4 * k4 + 1])), make_float2(b4.x, b4.y));
                    const float2 zb = add2_f32(make_float2(__uint_as_float(v[4 * k4 + 2]), __uint_as_float(v[4 * k4 + 3])), make_float2(b4.z, b4.w));
                    float2 la = mul2_f32(za, make_float2(0.2f, 0.2f)), lb = mul2_f32(zb, make_float2(0.2f, 0.2f));
                    la.x = fmaxf(la.x, za.x); la.y = fmaxf(la.y, za.y);
                    lb.x = fmaxf(lb.x, zb.x); lb.y = fmaxf(lb.y, zb.y);
                    const float2 ya = fma2_f32(la, make_float2(s4.x, s4.y), make_float2(t4.x, t4.y));
                    const float2 yb = fma2_f32(lb, make_float2(s4.z, s4.w), make_float2(t4.z, t4.w));
                    y[4 * k4 + 0] = ya.x; y[4 * k4 + 1] = ya.y; y[4 * k4 + 2] = yb.x; y[4 * k4 + 3] = yb.y;
                }
                unsigned char* rowp = stg + row * 128;
                const int swr = row & 7;
#pragma unroll
                for (int pc = 0; pc < 4; ++pc) {
                    uint4 o;
                    const float* yy = y + pc * 8;
                    o.x = pack2s(yy[0], yy[1]); o.y = pack2s(yy[2], yy[3]);
                    o.z = pack2s(yy[4], yy[5]); o.w = pack2s(yy[6], yy[7]);
                    *reinterpret_cast<uint4*>(rowp + (((g * 4 + pc) ^ swr) << 4)) = o;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(&tmem_empty[p], 0));
            if (elected && eg == 0) TR(4, it, 2);
            fence_proxy_async_smem();
            named_bar_sync(1 + eg, 128);
            if (elected && eg == 0) TR(4, it, 3);
            if (elected) {
                if (img < P.n_img) tma_store_4d(&map_out, stg, 0, w0, h, img);
                tma_store_commit();
                tma_store_wait_read();             // this group's staging buffer is free again for its next tile
            }
            if (elected && eg == 0) TR(4, it, 4);
            named_bar_sync(1 + eg, 128);
            if (elected && eg == 0) TR(4, it, 5);
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) tmem_dealloc2<C::TMEM_COLS>(tmem_base);
#else
    (void)P; (void)T;
    __trap();
#endif
}

// in: [n_img][H][W] fp32 on the device; stem parameters as HOST arrays (they travel in the kernel parameters);
// L = the conv2d_1 launch made by Net::make_launches (halo form)
void launch_stem_conv1(y3_context* ctx, const ConvLaunch& L, const float* in, const float* stem_w_host, const float* stem_bias_host,
                       const float* stem_scale_host, const float* stem_shift_host, int H, int W) {
    static bool attr[64] = {};
    if (!attr[ctx->device & 63]) {
        Y3_CUDA(cudaFuncSetAttribute(k_stem_conv1, cudaFuncAttributeMaxDynamicSharedMemorySize, SCfg::SMEM));
        attr[ctx->device & 63] = true;
    }
    const ConvArgs& A = L.args;
    StemArgs T;
    memcpy(T.w, stem_w_host, sizeof(T.w));
    memcpy(T.bias, stem_bias_host, sizeof(T.bias));
    memcpy(T.scale, stem_scale_host, sizeof(T.scale));
    memcpy(T.shift, stem_shift_host, sizeof(T.shift));
    T.in = in; T.H = H; T.W = W;
    T.dbg = nullptr;
#ifdef Y3_STEM_TRACE
    static long long* d_dbg = nullptr;
    if (!d_dbg) { cudaError_t e = cudaMalloc(&d_dbg, 65536); fprintf(stderr, "trace buffer %p (%s)\n", (void*)d_dbg, cudaGetErrorString(e)); }
    cudaMemsetAsync(d_dbg, 0, 65536, ctx->stream);
    T.dbg = d_dbg;
#endif
    const long long total = (long long)((A.n_img + 1) / 2) * A.tiles_x * A.Ho;
    const int pairs = (int)std::min<long long>(total, (long long)(ctx->sm_count / 2));
    k_stem_conv1<<<2 * pairs, SC_THREADS, SCfg::SMEM, ctx->stream>>>(L.map_b, L.map_out, L.args, T);
    Y3_LAUNCHED(ctx);
#ifdef Y3_STEM_TRACE
    static int traced = 0;
    if (++traced == 3) {
        long long hbuf[5 * 16 * 8];
        cudaStreamSynchronize(ctx->stream);
        cudaMemcpy(hbuf, d_dbg, sizeof(hbuf), cudaMemcpyDeviceToHost);
        long long base = hbuf[0];
        const char* roles[5] = {"mma", "stem0", "stem7", "epi0", "epi1"};
        for (int r = 0; r < 5; ++r)
            for (int st = 0; st < 16; ++st) {
                fprintf(stderr, "TRACE %-6s step %2d:", roles[r], 64 + st);
                for (int p = 0; p < 8; ++p) fprintf(stderr, " %7lld", hbuf[(r * 16 + st) * 8 + p] ? hbuf[(r * 16 + st) * 8 + p] - base : -1LL);
                fprintf(stderr, "\n");
            }
    }
#endif
}

}  // namespace y3
