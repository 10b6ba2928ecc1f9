// conv_tc2.cu - the 2-CTA (cta_group::2) variant of the implicit-GEMM convolution in conv_tc.cu.
//
// Why: ncu on the 1-CTA kernel (profiles/r1_ncu_conv_tc_*.txt) shows the 3x3 body layers bound by the
// L2 -> shared-memory fabric (11.6 TB/s, tensor pipe 35 %): a 128x128 CTA tile moves 32 KB per 64-deep
// K step = 64 FLOP/B.  Here two CTAs of a cluster (one TPC) form ONE UMMA of M = 256, N = BN2 (128 or
// 256): each CTA loads its own 128-pixel A patch and HALF of the weight tile, the tensor cores read
// both halves, so each CTA still moves <= 32 KB per K step but does 2x the math (128 FLOP/B at N=256).
//
//   * CTA rank r of pair-tile (mp, nt) owns M tile 2*mp + r (its own spatial patch) and weight rows
//     [nt*BN2 + r*BN2/2, +BN2/2).  TMA loads use the .cta_group::2 form and signal the LEADER's
//     (rank 0) full barrier; the leader's thread issues tcgen05.mma.cta_group::2 and commits with a
//     cluster multicast that frees the smem slot in BOTH CTAs and publishes the accumulator to both.
//   * accumulators: 2 x BN2 TMEM columns per CTA (double buffered), rows 0-127 = this CTA's pixels.
//   * epilogue (8 warps per CTA, two per TMEM lane quadrant) is the same fused
//     bias -> LeakyReLU(0.2) -> BN scale/shift -> (+ residual) -> bf16 -> TMA store as conv_tc.cu;
//     the non-leader's epilogue releases the accumulator with a remote mbarrier arrive.
#include "conv_tc.cuh"
#include "ptx.cuh"
#include <stdlib.h>

namespace y3 {
using namespace ptx;

static constexpr int CONV2_THREADS = 64 + 256;
static constexpr int TILE_M2 = 128;     // rows per CTA (the pair computes 256)

template <int BN2, int NSTG_>
struct Conv2Cfg {
    static constexpr int BK = 64;
    static constexpr int A_BYTES = TILE_M2 * BK * 2;
    static constexpr int B_BYTES = (BN2 / 2) * BK * 2;            // this CTA's half of the weight tile
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int NCHUNK = BN2 / 64;
    static constexpr int CHUNK_BYTES = TILE_M2 * 128;
    static constexpr int STG_BYTES = TILE_M2 * BN2 * 2;
    static constexpr int NSTG = NSTG_;                            // staging buffers (1: serialised store/residual)
    static constexpr int STAGES = (BN2 == 256) ? (NSTG_ == 1 ? 4 : 3) : 6;
    static constexpr int BAR_BYTES = (2 * STAGES + 8) * 8 + 16;
    static constexpr int SMEM = 1024 + STAGES * STAGE_BYTES + NSTG * STG_BYTES + BAR_BYTES;
    static constexpr uint32_t TMEM_COLS = 2 * BN2;
    static constexpr uint32_t SBO = 8 * BK * 2;
    static_assert(STAGE_BYTES % 1024 == 0, "operand tiles must stay 1024-B aligned");
    static_assert(SMEM <= 232448, "exceeds 227 KB of shared memory");
};

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

template <int BN2, int NSTG_>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CONV2_THREADS, 1)
k_conv_tc2(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a2,
          const __grid_constant__ CUtensorMap map_b,
           const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_res, const ConvArgs P) {
#if defined(__CUDA_ARCH_FEAT_SM100_ALL) || defined(__CUDA_ARCH_FEAT_SM101_ALL)
    using C = Conv2Cfg<BN2, NSTG_>;
    constexpr int BK = C::BK;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = align_smem_1024(smem_raw);
    unsigned char* stage_base = smem;
    unsigned char* stg_base = smem + C::STAGES * C::STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(stg_base + C::NSTG * C::STG_BYTES);
    uint64_t* full = bars;                       // used in the leader only (both CTAs' TMA land here)
    uint64_t* empty = bars + C::STAGES;          // one per CTA, multicast commit
    uint64_t* tmem_full = bars + 2 * C::STAGES;  // one per CTA, multicast commit
    uint64_t* tmem_empty = tmem_full + 2;        // leader only: 8 epilogue warps of each CTA arrive
    uint64_t* res_full = tmem_full + 4;
    uint64_t* stg_empty = tmem_full + 6;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 8);

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
            mbar_init(&tmem_empty[p], 16);       // 8 epilogue warps x 2 CTAs
            mbar_init(&res_full[p], 1);
            mbar_init(&stg_empty[p], 1);
        }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc2<C::TMEM_COLS>(tmem_slot);     // same warp id in both CTAs (collective)
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_launch_dependents();
    griddep_wait();              // activations of the previous layer are complete and visible from here on

    const int rows = P.BH * P.BW;
    const int k_iters = P.taps * P.kchunks;
    const int m_tiles = P.tiles_per_img * P.n_img;
    const int m_pairs = (m_tiles + 1) >> 1;
    const int total_pt = m_pairs * P.n_tiles_n;                // pair tiles

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (each CTA)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int t = pair; t < total_pt; t += n_pairs, ++it) {
                const int nt = t % P.n_tiles_n;
                const int mt = 2 * (t / P.n_tiles_n) + (int)rank;
                const int img = mt / P.tiles_per_img;
                const int r = mt - img * P.tiles_per_img;
                const int ty = r / P.tiles_x;
                const int x0 = (r - ty * P.tiles_x) * P.BW;
                const int y0 = ty * P.BH;
                const int n0 = nt * BN2;
                int im_w = 0, im_h = 0, im_n = 0;
                if (P.im2col) {                               // x0 = first flattened output pixel of this tile
                    const int per = P.im_ho * P.im_wo;
                    im_n = x0 / per;
                    const int rem = x0 - im_n * per;
                    const int oh_ = rem / P.im_wo;
                    im_h = oh_ * P.im_stride + P.im_lower;
                    im_w = (rem - oh_ * P.im_wo) * P.im_stride + P.im_lower;
                }
                const int p = (C::NSTG == 2) ? (it & 1) : 0;
                const uint32_t use = (C::NSTG == 2) ? (uint32_t)(it >> 1) : (uint32_t)it;
                if (P.has_res) {
                    mbar_wait(&stg_empty[p], (use & 1u) ^ 1u);
                    mbar_expect_tx(&res_full[p], (uint32_t)(rows * BN2 * 2));
                    for (int ch = 0; ch < C::NCHUNK; ++ch)
                        tma_load_4d(stg_base + p * C::STG_BYTES + ch * C::CHUNK_BYTES, &map_res, &res_full[p], n0 + ch * 64, x0,
                                    y0, img);
                }
                for (int tap = 0; tap < P.taps; ++tap) {
                    const int kh = tap / P.kwn;
                    const int kw = tap - kh * P.kwn;
                    for (int kc = 0; kc < P.kchunks; ++kc) {
                        mbar_wait(&empty[stage], phase ^ 1u);
                        unsigned char* sa = stage_base + stage * C::STAGE_BYTES;
                        const uint32_t lead_full = mapa_u32(&full[stage], 0);
                        if (rank == 0) mbar_expect_tx(&full[stage], (uint32_t)(2 * (rows * BK * 2 + C::B_BYTES)));
                        if (P.im2col) {
                            tma2_load_im2col_4d(sa, &map_a, lead_full, kc * BK, im_w, im_h, im_n, (uint16_t)kw, (uint16_t)kh);
                        } else if (P.stride == 1) {
                            if (kc < P.k_split) tma2_load_4d(sa, &map_a, lead_full, kc * BK, x0 + kw - P.pad, y0 + kh - P.pad, img);
                            else tma2_load_4d(sa, &map_a2, lead_full, (kc - P.k_split) * BK, x0 + kw - P.pad, y0 + kh - P.pad, img);
                        }
                        else
                            tma2_load_5d(sa, &map_a, lead_full, (kw & 1) * P.a_cpitch + kc * BK, x0 + (kw >> 1), kh & 1,
                                         y0 + (kh >> 1), img);
                        tma2_load_2d(sa + C::A_BYTES, &map_b, lead_full, tap * P.cin + kc * BK, n0 + (int)rank * (BN2 / 2));
                        if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
            }
            // producer tail: do not let this CTA exit while multicast commits from the leader may still
            // arrive on its empty barriers
            for (int s = 0; s < C::STAGES; ++s) {
                mbar_wait(&empty[stage], phase ^ 1u);
                if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer: leader CTA, one thread
        if (rank == 0 && lane == 0) {
            const uint32_t idesc1 = P.in_f16 ? make_idesc_f16(256, BN2) : make_idesc_bf16(256, BN2);
            const uint32_t idesc2 = P.in2_f16 ? make_idesc_f16(256, BN2) : make_idesc_bf16(256, BN2);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int t = pair; t < total_pt; t += n_pairs, ++it) {
                const int p = it & 1;
                const uint32_t use = (uint32_t)(it >> 1);
                mbar_wait(&tmem_empty[p], (use & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(p * BN2);
                int kc = 0;
                for (int ki = 0; ki < k_iters; ++ki) {
                    const uint32_t idesc = kc < P.k_split ? idesc1 : idesc2;
                    if (++kc == P.kchunks) kc = 0;
                    mbar_wait(&full[stage], phase);
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
        }
    } else {
        // ------------------------------------------------------------ epilogue: warps 2..9 of each CTA
        const int ew = warp - 2;
        const int q = warp & 3;                    // TMEM lane quadrant
        const int half = ew >> 2;                  // which half of the column groups
        const int row = q * 32 + lane;
        const int by = row / P.BW;
        const int bx = row - by * P.BW;
        int it = 0;
        for (int t = pair; t < total_pt; t += n_pairs, ++it) {
            const int nt = t % P.n_tiles_n;
            const int mt = 2 * (t / P.n_tiles_n) + (int)rank;
            const int img = mt / P.tiles_per_img;
            const int r = mt - img * P.tiles_per_img;
            const int ty = r / P.tiles_x;
            const int x0 = (r - ty * P.tiles_x) * P.BW;
            const int y0 = ty * P.BH;
            const int n0 = nt * BN2;
            const int p = it & 1;
            const uint32_t use = (uint32_t)(it >> 1);
            const int sp = (C::NSTG == 2) ? p : 0;
            const uint32_t suse = (C::NSTG == 2) ? use : (uint32_t)it;
            mbar_wait(&tmem_full[p], use & 1u);
            tc_fence_after();
            if (P.has_res) mbar_wait(&res_full[sp], suse & 1u);
            unsigned char* stg = stg_base + sp * C::STG_BYTES;
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(p * BN2);
            const bool pix_ok = (mt < m_tiles) && (row < rows) && (y0 + by < P.Ho) && (x0 + bx < P.Wo);
#pragma unroll 1
            for (int g = half; g < BN2 / 32; g += 2) {
                uint32_t v[32];
                tmem_ld_32x32(t_addr + (uint32_t)(g * 32), v);
                tmem_ld_wait();
                const int c0 = n0 + g * 32;
                float y[32];
                if (P.linear) epilogue_math_32<true>(v, P.bias + c0, P.scale + c0, P.shift + c0, y);
                else epilogue_math_32<false>(v, P.bias + c0, P.scale + c0, P.shift + c0, y);
                if (P.out_f32) {
                    if (pix_ok) {
                        const long long pix = ((long long)img * P.Ho + (y0 + by)) * P.Wo + (x0 + bx);
                        float4* dst = reinterpret_cast<float4*>(P.out32 + pix * P.out32_pitch + c0);
#pragma unroll
                        for (int k4 = 0; k4 < 8; ++k4)
                            dst[k4] = make_float4(y[4 * k4], y[4 * k4 + 1], y[4 * k4 + 2], y[4 * k4 + 3]);
                    }
                } else {
                    const int ch = g >> 1;
                    const int piece0 = (g & 1) * 4;
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
            // accumulator drained: release it to the leader's MMA thread (remote arrive from rank 1)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(&tmem_empty[p], 0));
            if (!P.out_f32) {
                fence_proxy_async_smem();
                // drain the previous tile's store (other staging buffer) before anyone may pass the barrier
                if (warp == 2 && lane == 0 && !P.has_res && C::NSTG == 2) tma_store_wait_read();
                named_bar_sync(1, 256);
                if (warp == 2 && lane == 0) {
                    if (mt < m_tiles)
                        for (int ch = 0; ch < C::NCHUNK; ++ch)
                            tma_store_4d(&map_out, stg + ch * C::CHUNK_BYTES, n0 + ch * 64, x0, y0, img);
                    tma_store_commit();
                    if (P.has_res || C::NSTG == 1) {
                        tma_store_wait_read();
                        if (P.has_res) mbar_arrive(&stg_empty[sp]);
                    }                                         // (no residual: drained lazily before the next barrier)
                }
                if (C::NSTG == 1) named_bar_sync(2, 256);     // single staging buffer: wait until it was read
            }
        }
    }
    if (warp == 2 && lane == 0) tma_store_wait_read();
    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) tmem_dealloc2<C::TMEM_COLS>(tmem_base);
#else
    (void)P;
    __trap();
#endif
}

template <int BN2, int NSTG_>
static void launch2_t(y3_context* ctx, const ConvLaunch& L) {
    using C = Conv2Cfg<BN2, NSTG_>;
    static bool attr[64] = {};
    if (!attr[ctx->device & 63]) {
        Y3_CUDA(cudaFuncSetAttribute(k_conv_tc2<BN2, NSTG_>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        attr[ctx->device & 63] = true;
    }
    launch_pdl(k_conv_tc2<BN2, NSTG_>, L.grid, CONV2_THREADS, C::SMEM, ctx->stream, L.map_a, L.map_a2, L.map_b, L.map_out, L.map_res, L.args);
    Y3_LAUNCHED(ctx);
}

void launch_conv2(y3_context* ctx, const ConvLaunch& L) {
    if (launch_conv2h(ctx, L)) return;
    static const int stg256 = getenv("Y3_CONV2_STG") ? atoi(getenv("Y3_CONV2_STG")) : 2;
    if (L.bn == 256 && stg256 == 1) launch2_t<256, 1>(ctx, L);
    else if (L.bn == 256) launch2_t<256, 2>(ctx, L);
    else if (L.bn == 128) launch2_t<128, 2>(ctx, L);
    else fail(Y3_ERR_UNSUPPORTED, "no 2-CTA conv kernel for BN2=%d", L.bn);
}

}  // namespace y3
