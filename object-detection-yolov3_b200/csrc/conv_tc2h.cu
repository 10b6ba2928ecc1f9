// conv_tc2h.cu - 2-CTA (cta_group::2) implicit-GEMM convolution, "half-staged" epilogue.
//
// Same math and pairing as conv_tc2.cu (UMMA M = 256, N = BN2; each CTA loads its own 128-pixel A tile and
// half of the weight tile).  What changes is the back end, to buy pipeline depth:
//   * the output tile is staged 128 columns (32 KB) at a time through TWO half buffers instead of two
//     whole-tile buffers, which leaves room for 5 operand stages at N = 256 (3 before): ~160 KB of loads in
//     flight per CTA instead of 96 KB - the N = 256 layers were latency-limited (ncu: tensor 50 %, fabric 72 %);
//   * a dedicated staging-manager warp (warp 10) issues the residual prefetch (TMA load of the block input
//     into the half buffer) and the TMA stores, and waits for the stores to drain, so the eight epilogue warps
//     never block on the TMA: they only wait "buffer ready" and signal "half staged" through mbarriers.
//
//   warp 0      TMA producer of the (A, B) operand ring
//   warp 1      tcgen05.mma issuer (leader CTA, one thread)
//   warps 2-9   epilogue: TMEM -> bias/LeakyReLU/BN(+residual) -> bf16 -> swizzled half buffer
//   warp 10     staging manager: residual loads, output stores, buffer recycling
#include "conv_tc.cuh"
#include "ptx.cuh"
#include <stdlib.h>

namespace y3 {
using namespace ptx;

static constexpr int CONV2H_THREADS = 64 + 256 + 32;

template <int BN2>
struct Conv2hCfg {
    static constexpr int BK = 64;
    static constexpr int A_BYTES = 128 * BK * 2;
    static constexpr int B_BYTES = (BN2 / 2) * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int NH = BN2 / 128;                  // 128-column halves per tile
    static constexpr int CHUNK_BYTES = 128 * 128;         // 128 rows x 64 channels
    static constexpr int HS_BYTES = 2 * CHUNK_BYTES;      // one half buffer (128 rows x 128 channels bf16)
    static constexpr int STAGES = (BN2 == 256) ? 5 : 6;
    static constexpr int NBARS = 2 * STAGES + 12;
    static constexpr int SMEM = 1024 + STAGES * STAGE_BYTES + 2 * HS_BYTES + NBARS * 8 + 16;
    static constexpr uint32_t TMEM_COLS = 2 * BN2;
    static constexpr uint32_t SBO = 8 * BK * 2;
    static_assert(STAGE_BYTES % 1024 == 0, "operand tiles must stay 1024-B aligned");
    static_assert(SMEM <= 232448, "exceeds 227 KB of shared memory");
};

struct TileCoord { int n0, x0, y0, img, mt; };

#ifdef Y3_CONV_TRACE
// measurement build only: where CTA 0's MMA thread, producer and first epilogue warp spend their cycles
__device__ long long g_conv_trace[16];
#define CT_BEGIN(v) const long long v = clock64()
#define CT_ADD(slot, v) do { if (blockIdx.x == 0) g_conv_trace[slot] += clock64() - (v); } while (0)
#else
#define CT_BEGIN(v) do { } while (0)
#define CT_ADD(slot, v) do { } while (0)
#endif

__device__ __forceinline__ TileCoord tile_coord(const ConvArgs& P, int t, int rank, int bn2) {
    TileCoord c;
    const int nt = t % P.n_tiles_n;
    c.mt = 2 * (t / P.n_tiles_n) + rank;
    c.img = c.mt / P.tiles_per_img;
    const int r = c.mt - c.img * P.tiles_per_img;
    const int ty = r / P.tiles_x;
    c.x0 = (r - ty * P.tiles_x) * P.BW;
    c.y0 = ty * P.BH;
    c.n0 = nt * bn2;
    return c;
}

__device__ __forceinline__ uint32_t pack2h(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

template <int BN2>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CONV2H_THREADS, 1)
k_conv_tc2h(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a2,
            const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_out,
            const __grid_constant__ CUtensorMap map_res, const ConvArgs P) {
#if defined(__CUDA_ARCH_FEAT_SM100_ALL) || defined(__CUDA_ARCH_FEAT_SM101_ALL)
    using C = Conv2hCfg<BN2>;
    constexpr int BK = C::BK;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = align_smem_1024(smem_raw);
    unsigned char* stage_base = smem;
    unsigned char* hs_base = smem + C::STAGES * C::STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(hs_base + 2 * C::HS_BYTES);
    uint64_t* full = bars;                        // leader only
    uint64_t* empty = bars + C::STAGES;
    uint64_t* tmem_full = bars + 2 * C::STAGES;   // [2]
    uint64_t* tmem_empty = tmem_full + 2;         // [2] leader only, 16 arrivals
    uint64_t* res_full = tmem_full + 4;           // [2] residual half landed (tx)
    uint64_t* buf_free = tmem_full + 6;           // [2] manager -> epilogue: half buffer may be overwritten
    uint64_t* staged = tmem_full + 8;             // [2] epilogue (8 warps) -> manager: half is in the buffer
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 12);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1;
    const int n_pairs = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_a);
        if (P.k_split < P.kchunks) prefetch_tmap(&map_a2);
        prefetch_tmap(&map_b);
        if (!P.out_f32) prefetch_tmap(&map_out);
        if (P.has_res) prefetch_tmap(&map_res);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int p = 0; p < 2; ++p) {
            mbar_init(&tmem_full[p], 1);
            mbar_init(&tmem_empty[p], 16);
            mbar_init(&res_full[p], 1);
            mbar_init(&buf_free[p], 1);
            mbar_init(&staged[p], 8);
        }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc2<C::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_launch_dependents();
    griddep_wait();              // activations of the previous layer are complete and visible from here on

    const int rows = P.BH * P.BW;
    const int k_iters = P.taps * P.kchunks;
    const int m_tiles = P.tiles_per_img * P.n_img;
    const int total_pt = ((m_tiles + 1) >> 1) * P.n_tiles_n;

    if (warp == 0) {
        // ------------------------------------------------------------ operand producer (each CTA)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = pair; t < total_pt; t += n_pairs) {
                const TileCoord c = tile_coord(P, t, (int)rank, BN2);
                int im_w = 0, im_h = 0, im_n = 0;
                if (P.im2col) {
                    const int per = P.im_ho * P.im_wo;
                    im_n = c.x0 / per;
                    const int rem = c.x0 - im_n * per;
                    const int oh_ = rem / P.im_wo;
                    im_h = oh_ * P.im_stride + P.im_lower;
                    im_w = (rem - oh_ * P.im_wo) * P.im_stride + P.im_lower;
                }
                for (int tap = 0; tap < P.taps; ++tap) {
                    const int kh = tap / P.kwn;
                    const int kw = tap - kh * P.kwn;
                    for (int kc = 0; kc < P.kchunks; ++kc) {
                        CT_BEGIN(c0);
                        mbar_wait(&empty[stage], phase ^ 1u);
                        CT_ADD(0, c0);
                        unsigned char* sa = stage_base + stage * C::STAGE_BYTES;
                        const uint32_t lead_full = mapa_u32(&full[stage], 0);
                        if (rank == 0) mbar_expect_tx(&full[stage], (uint32_t)(2 * (rows * BK * 2 + C::B_BYTES)));
                        // weights first: every pair asks for the same slice at about the same time (the L2 merges those
                        // requests); measured -1.6 % on the 128->256 @64^2 class against activations-first
                        tma2_load_2d(sa + C::A_BYTES, &map_b, lead_full, tap * P.cin + kc * BK, c.n0 + (int)rank * (BN2 / 2));
                        if (P.im2col) {
                            tma2_load_im2col_4d(sa, &map_a, lead_full, kc * BK, im_w, im_h, im_n, (uint16_t)kw, (uint16_t)kh);
                        } else if (P.stride == 1) {
                            if (kc < P.k_split) tma2_load_4d(sa, &map_a, lead_full, kc * BK, c.x0 + kw - P.pad, c.y0 + kh - P.pad, c.img);
                            else tma2_load_4d(sa, &map_a2, lead_full, (kc - P.k_split) * BK, c.x0 + kw - P.pad, c.y0 + kh - P.pad, c.img);
                        } else {
                            tma2_load_5d(sa, &map_a, lead_full, (kw & 1) * P.a_cpitch + kc * BK, c.x0 + (kw >> 1), kh & 1, c.y0 + (kh >> 1), c.img);
                        }
                        if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
            }
            for (int s = 0; s < C::STAGES; ++s) {      // tail: multicast commits may still target this CTA
                mbar_wait(&empty[stage], phase ^ 1u);
                if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (leader CTA, one thread)
        if (rank == 0 && lane == 0) {
            const uint32_t idesc1 = P.in_f16 ? make_idesc_f16(256, BN2) : make_idesc_bf16(256, BN2);
            const uint32_t idesc2 = P.in2_f16 ? make_idesc_f16(256, BN2) : make_idesc_bf16(256, BN2);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            CT_BEGIN(c_all);
            for (int t = pair; t < total_pt; t += n_pairs, ++it) {
                const int p = it & 1;
                const uint32_t use = (uint32_t)(it >> 1);
                CT_BEGIN(c1);
                mbar_wait(&tmem_empty[p], (use & 1u) ^ 1u);
                CT_ADD(1, c1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(p * BN2);
                int kc = 0;
                for (int ki = 0; ki < k_iters; ++ki) {
                    const uint32_t idesc = kc < P.k_split ? idesc1 : idesc2;
                    if (++kc == P.kchunks) kc = 0;
                    CT_BEGIN(c2);
                    mbar_wait(&full[stage], phase);
                    CT_ADD(2, c2);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(stage_base + stage * C::STAGE_BYTES);
                    const uint64_t adesc = make_smem_desc(a_addr, C::SBO, SWZ_128B);
                    const uint64_t bdesc = make_smem_desc(a_addr + C::A_BYTES, C::SBO, SWZ_128B);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        umma2_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (uint32_t)((ki | k) != 0));
                    umma2_commit_mc(&empty[stage], 3);
                    if (ki == k_iters - 1) umma2_commit_mc(&tmem_full[p], 3);
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
                }
            }
            CT_ADD(3, c_all);
#ifdef Y3_CONV_TRACE
            if (blockIdx.x == 0) g_conv_trace[4] += it;
#endif
        }
    } else if (warp == 10) {
        // ------------------------------------------------------------ staging manager (each CTA, one thread)
        // half hc (global half counter) uses buffer hc & 1.  Order per half: [residual prefetch for hc+1 into
        // the buffer whose store was drained one iteration ago] -> wait "staged(hc)" -> store -> drain -> free.
        if (lane == 0 && !P.out_f32) {
            const int my_tiles = (total_pt - pair + n_pairs - 1) / n_pairs;
            const int n_half = my_tiles * C::NH;
            auto half_coord = [&](int hc, TileCoord* c, int* h) {
                const int it = hc / C::NH;
                *h = hc - it * C::NH;
                *c = tile_coord(P, pair + it * n_pairs, (int)rank, BN2);
            };
            auto prefetch_res = [&](int hc) {
                TileCoord c; int h;
                half_coord(hc, &c, &h);
                const int buf = hc & 1;
                unsigned char* dst = hs_base + buf * C::HS_BYTES;
                mbar_expect_tx(&res_full[buf], (uint32_t)(rows * 128 * 2));
                tma_load_4d(dst, &map_res, &res_full[buf], c.n0 + h * 128, c.x0, c.y0, c.img);
                tma_load_4d(dst + C::CHUNK_BYTES, &map_res, &res_full[buf], c.n0 + h * 128 + 64, c.x0, c.y0, c.img);
            };
            if (P.has_res) {                       // both buffers are free at the start
                if (n_half > 0) prefetch_res(0);
                if (n_half > 1) prefetch_res(1);
            }
            for (int hc = 0; hc < n_half; ++hc) {
                const int buf = hc & 1;
                const uint32_t use = (uint32_t)(hc >> 1);
                TileCoord c; int h;
                half_coord(hc, &c, &h);
                mbar_wait(&staged[buf], use & 1u);                 // the 8 epilogue warps wrote + fenced this half
                unsigned char* src = hs_base + buf * C::HS_BYTES;
                if (c.mt < m_tiles) {
                    tma_store_4d(&map_out, src, c.n0 + h * 128, c.x0, c.y0, c.img);
                    tma_store_4d(&map_out, src + C::CHUNK_BYTES, c.n0 + h * 128 + 64, c.x0, c.y0, c.img);
                }
                tma_store_commit();
                tma_store_wait_read();                              // only this thread blocks on the drain
                if (hc + 2 < n_half) {
                    if (P.has_res) prefetch_res(hc + 2);            // refill the drained buffer (res_full => free)
                    else mbar_arrive(&buf_free[buf]);
                }
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue: warps 2..9 of each CTA
        const int ew = warp - 2;
        const int q = warp & 3;                    // TMEM lane quadrant
        const int gsel = ew >> 2;                  // which column groups of a half (0: g 0,2  1: g 1,3)
        const int row = q * 32 + lane;
        const int by = row / P.BW;
        const int bx = row - by * P.BW;
        int it = 0, hc = 0;
        for (int t = pair; t < total_pt; t += n_pairs, ++it) {
            const TileCoord c = tile_coord(P, t, (int)rank, BN2);
            const int p = it & 1;
            const uint32_t use = (uint32_t)(it >> 1);
            CT_BEGIN(c5);
            mbar_wait(&tmem_full[p], use & 1u);
            if (warp == 2 && lane == 0) CT_ADD(5, c5);
            CT_BEGIN(c6);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(p * BN2);
            const bool pix_ok = (c.mt < m_tiles) && (row < rows) && (c.y0 + by < P.Ho) && (c.x0 + bx < P.Wo);
            for (int h = 0; h < C::NH; ++h, ++hc) {
                const int buf = hc & 1;
                const uint32_t huse = (uint32_t)(hc >> 1);
                unsigned char* stg = hs_base + buf * C::HS_BYTES;
                CT_BEGIN(c7);
                if (!P.out_f32) {
                    if (P.has_res) mbar_wait(&res_full[buf], huse & 1u);          // residual landed => buffer is ours
                    else if (hc >= 2) mbar_wait(&buf_free[buf], (huse & 1u) ^ 1u); // store of half hc-2 drained
                }
                if (warp == 2 && lane == 0) CT_ADD(7, c7);
#pragma unroll 1
                for (int g2 = gsel; g2 < 4; g2 += 2) {
                    const int g = h * 4 + g2;
                    uint32_t v[32];
                    tmem_ld_32x32(t_addr + (uint32_t)(g * 32), v);
                    tmem_ld_wait();
                    const int c0 = c.n0 + g * 32;
                    float y[32];
                    if (P.linear) epilogue_math_32<true>(v, P.bias + c0, P.scale + c0, P.shift + c0, y);
                    else epilogue_math_32<false>(v, P.bias + c0, P.scale + c0, P.shift + c0, y);
                    if (P.out_f32) {
                        if (pix_ok) {
                            const long long pix = ((long long)c.img * P.Ho + (c.y0 + by)) * P.Wo + (c.x0 + bx);
                            float4* dst = reinterpret_cast<float4*>(P.out32 + pix * P.out32_pitch + c0);
#pragma unroll
                            for (int k4 = 0; k4 < 8; ++k4)
                                dst[k4] = make_float4(y[4 * k4], y[4 * k4 + 1], y[4 * k4 + 2], y[4 * k4 + 3]);
                        }
                    } else {
                        const int ch = g2 >> 1;
                        const int piece0 = (g2 & 1) * 4;
                        unsigned char* rowp = stg + ch * C::CHUNK_BYTES + row * 128;
                        const int sw = row & 7;
                        if (P.has_res) {                     // block input: all four pieces requested first, one round trip
#pragma unroll
                            for (int pc = 0; pc < 4; ++pc) {
                                const uint4 x = lds128(rowp + (((piece0 + pc) ^ sw) << 4));
                                float* yy = y + pc * 8;
                                yy[0] += __uint_as_float(x.x << 16); yy[1] += __uint_as_float(x.x & 0xffff0000u);
                                yy[2] += __uint_as_float(x.y << 16); yy[3] += __uint_as_float(x.y & 0xffff0000u);
                                yy[4] += __uint_as_float(x.z << 16); yy[5] += __uint_as_float(x.z & 0xffff0000u);
                                yy[6] += __uint_as_float(x.w << 16); yy[7] += __uint_as_float(x.w & 0xffff0000u);
                            }
                        }
                        if (P.out_f16) stage_row_32<true>(y, rowp, piece0, sw);
                        else stage_row_32<false>(y, rowp, piece0, sw);
                    }
                }
                if (h == C::NH - 1) {              // accumulator fully drained: release it to the leader's MMA thread
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(mapa_u32(&tmem_empty[p], 0));
                    if (warp == 2 && lane == 0) CT_ADD(6, c6);
                }
                if (!P.out_f32) {
                    fence_proxy_async_smem();      // generic-proxy writes -> visible to the TMA store
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&staged[buf]);
                }
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) tmem_dealloc2<C::TMEM_COLS>(tmem_base);
#else
    (void)P;
    __trap();
#endif
}

template <int BN2>
static void launch2h_t(y3_context* ctx, const ConvLaunch& L) {
    using C = Conv2hCfg<BN2>;
    static bool attr[64] = {};
    if (!attr[ctx->device & 63]) {
        Y3_CUDA(cudaFuncSetAttribute(k_conv_tc2h<BN2>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        attr[ctx->device & 63] = true;
    }
#ifdef Y3_CONV_TRACE
    long long zero[16] = {};
    cudaMemcpyToSymbol(g_conv_trace, zero, sizeof(zero));
#endif
    launch_pdl(k_conv_tc2h<BN2>, L.grid, CONV2H_THREADS, C::SMEM, ctx->stream, L.map_a, L.map_a2, L.map_b, L.map_out, L.map_res, L.args);
    Y3_LAUNCHED(ctx);
#ifdef Y3_CONV_TRACE
    long long tr[16];
    cudaStreamSynchronize(ctx->stream);
    cudaMemcpyFromSymbol(tr, g_conv_trace, sizeof(tr));
    fprintf(stderr, "CONVTRACE cin %d taps %d kchunks %d ntn %d tiles %lld: mma total %lld = wait_tmem_empty %lld + wait_full %lld + issue %lld | producer wait_empty %lld | epi wait_tmem_full %lld, hold_acc %lld (of which wait_buffer %lld)\n",
            L.args.cin, L.args.taps, L.args.kchunks, L.args.n_tiles_n, tr[4], tr[3], tr[1], tr[2], tr[3] - tr[1] - tr[2], tr[0], tr[5], tr[6], tr[7]);
#endif
}

bool launch_conv2h(y3_context* ctx, const ConvLaunch& L) {
    static const bool enabled = getenv("Y3_CONV2_OLD") == nullptr;
    // measured per layer class (profiles/r1_layers_*): the deeper operand ring pays off for the long 3x3
    // mainloops at N = 256; N = 128 tiles and the 4-16-step 1x1 layers are faster with whole-tile staging
    if (!enabled || L.bn != 256 || L.args.taps == 1) return false;
    launch2h_t<256>(ctx, L);
    return true;
}

}  // namespace y3
