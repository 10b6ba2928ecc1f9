// conv_halo.cu - weights-stationary, halo-row variant of the 3x3 stride-1 convolution
// (YoloV3.conv_layer, reference model.py:29-39) for the shallow layers (Cin 32 / 64), which are bound by
// the L2 -> shared-memory fabric in the im2col kernels: there every 128-pixel M tile re-fetches its
// input 9 times (once per tap) and the whole weight tile once, 4.2 GB over the fabric for 0.67 GB of
// tensors (profiles/r1_ncu_conv_tc_64x32_*.txt).
//
// Here
//   * the 9 weight taps of the layer (this CTA's half of the Cout rows) are loaded ONCE per CTA and stay
//     in shared memory for the whole launch;
//   * an M tile is 128 consecutive pixels of one output row.  The CTA walks DOWN a column of such tiles
//     and keeps a ring of input rows (130 pixels = tile + 1 halo pixel each side, zero-filled by TMA at
//     the image border = TF "SAME" padding) in shared memory: every new output row costs ONE new input
//     row, and the 9 taps are addressed by the UMMA descriptor alone - ring slot = dy, start address
//     shifted by dx rows of the K-major swizzled tile.  (The swizzle is a function of the absolute smem
//     address, so a row-shifted start reads exactly what TMA wrote: tests/probe_umma_rowshift.py.)
//   * two CTAs form one cta_group::2 UMMA (M = 256: rank r works on image 2*ip + r, same column, same
//     rows, so both rings stay in lock step), N = Cout, each CTA holds half of the weight rows.
//   * epilogue identical to conv_tc2.cu: bias -> LeakyReLU(0.2) -> BN scale/shift -> + block input ->
//     bf16 -> 128B-swizzled staging -> TMA store; residual tile prefetched by TMA into the staging buffer.
#include "conv_tc.cuh"
#include "ptx.cuh"
#include <stdlib.h>
#include <algorithm>

namespace y3 {
using namespace ptx;

static constexpr int HALO_THREADS = 64 + 256;

template <int CIN, int COUT, int STRIDE>
struct HaloCfg {
    static constexpr int ROWB = CIN * 2;                              // bytes of one pixel = one K-major row
    static constexpr int HPX = (STRIDE == 1) ? 130 : 128;             // pixels per ring slot
    static constexpr int SLOT = ((HPX * ROWB + 1023) / 1024) * 1024;
    // stride 1: ring of input rows (3 live + prefetch); stride 2: ring of per-tap im2col tiles
    static constexpr int S = (STRIDE == 1) ? ((CIN == 32) ? 6 : 5) : ((CIN == 32) ? 7 : 5);
    static constexpr int WTAP = (COUT / 2) * ROWB;                    // this CTA's half of one tap
    static constexpr int W_BYTES = 9 * WTAP;
    static constexpr int NCHUNK = COUT / 64;
    static constexpr int CHUNK_BYTES = 128 * 128;
    static constexpr int STG_BYTES = 128 * COUT * 2;
    static constexpr int BAR_BYTES = (2 * S + 10) * 8 + 16;
    static constexpr int PAR_BYTES = 3 * COUT * 4;                    // bias | scale | shift
    static constexpr int SMEM = 1024 + W_BYTES + S * SLOT + 2 * STG_BYTES + PAR_BYTES + BAR_BYTES;
    static constexpr int CTAS_PER_SM = (CIN == 32) ? 2 : 1;
    static constexpr uint32_t TMEM_COLS = 2 * COUT;
    static constexpr uint32_t SWZ = (CIN == 64) ? (uint32_t)SWZ_128B : (uint32_t)SWZ_64B;
    static constexpr uint32_t SBO = 8 * ROWB;
    static_assert(WTAP % 1024 == 0, "weight taps must stay 1024-B aligned");
    static_assert(SMEM * CTAS_PER_SM <= 232448 - 1024 * CTAS_PER_SM, "exceeds 227 KB of shared memory");
};

__device__ __forceinline__ uint32_t pack2h(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// P.tiles_x = segments per row, P.Ho/P.Wo = output grid, P.n_img = images of this launch
template <int CIN, int COUT, int STRIDE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(HALO_THREADS, HaloCfg<CIN, COUT, STRIDE>::CTAS_PER_SM)
k_conv_halo(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
            const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_res, const ConvArgs P) {
#if defined(__CUDA_ARCH_FEAT_SM100_ALL) || defined(__CUDA_ARCH_FEAT_SM101_ALL)
    using C = HaloCfg<CIN, COUT, STRIDE>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = align_smem_1024(smem_raw);
    unsigned char* w_base = smem;
    unsigned char* ring = smem + C::W_BYTES;
    unsigned char* stg_base = ring + C::S * C::SLOT;
    float* s_par = reinterpret_cast<float*>(stg_base + 2 * C::STG_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(stg_base + 2 * C::STG_BYTES + C::PAR_BYTES);
    uint64_t* full = bars;                       // leader only: both CTAs' row loads land here
    uint64_t* empty = bars + C::S;               // per CTA, multicast commit
    uint64_t* tmem_full = bars + 2 * C::S;       // per CTA, multicast commit
    uint64_t* tmem_empty = tmem_full + 2;        // leader only
    uint64_t* res_full = tmem_full + 4;
    uint64_t* stg_empty = tmem_full + 6;
    uint64_t* w_full = tmem_full + 8;            // leader only: both halves of the weights
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 10);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1;
    const int n_pairs = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_a);
        prefetch_tmap(&map_b);
        prefetch_tmap(&map_out);
        if (P.has_res) prefetch_tmap(&map_res);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < C::S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int p = 0; p < 2; ++p) {
            mbar_init(&tmem_full[p], 1);
            mbar_init(&tmem_empty[p], 16);       // 8 epilogue warps x 2 CTAs
            mbar_init(&res_full[p], 1);
            mbar_init(&stg_empty[p], 1);
        }
        mbar_init(w_full, 1);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc2<C::TMEM_COLS>(tmem_slot);
    for (int i = threadIdx.x; i < COUT; i += HALO_THREADS) {
        s_par[i] = P.bias[i]; s_par[COUT + i] = P.scale[i]; s_par[2 * COUT + i] = P.shift[i];
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_launch_dependents();
    if (warp == 0 && lane == 0) {
        // stationary weights: 9 taps x this CTA's half of the Cout rows; they do not depend on the previous layer,
        // so they are requested before the grid-dependency wait
        const uint32_t lead_w = mapa_u32(w_full, 0);
        if (rank == 0) mbar_expect_tx(w_full, (uint32_t)(2 * C::W_BYTES));
        for (int tap = 0; tap < 9; ++tap)
            tma2_load_2d(w_base + tap * C::WTAP, &map_b, lead_w, tap * CIN, (int)rank * (COUT / 2));
    }
    griddep_wait();              // activations of the previous layer are complete and visible from here on

    // pair-tiles: (image pair ip, segment seg, output row h) flattened with h fastest; this pair's range
    const int Ho = P.Ho;
    const int img_pairs = (P.n_img + 1) >> 1;
    const long long total = (long long)img_pairs * P.tiles_x * Ho;
    const int t_begin = (int)(total * pair / n_pairs);
    const int t_end = (int)(total * (pair + 1) / n_pairs);

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (each CTA)
        if (lane == 0) {
            int k = 0;                                  // running input-row load index (ring position)
            int it = 0;                                 // running tile index
            if (STRIDE == 2) {
                // stride 2: the A operand of every tap is an im2col-mode load of 128 output pixels (padding 0 before /
                // 1 after and the traversal stride live in the map); only the weights are stationary
                for (int t = t_begin; t < t_end; ++t, ++it) {
                    const int col = t / Ho;
                    const int h = t - col * Ho;
                    const int ip = col / P.tiles_x;
                    const int w0 = (col - ip * P.tiles_x) * 128;
                    const int img = 2 * ip + (int)rank;
                    if (P.has_res) {
                        const int p = it & 1;
                        const uint32_t use = (uint32_t)(it >> 1);
                        mbar_wait(&stg_empty[p], (use & 1u) ^ 1u);
                        mbar_expect_tx(&res_full[p], (uint32_t)C::STG_BYTES);
                        for (int ch = 0; ch < C::NCHUNK; ++ch)
                            tma_load_4d(stg_base + p * C::STG_BYTES + ch * C::CHUNK_BYTES, &map_res, &res_full[p], ch * 64, w0, h, img);
                    }
                    for (int tap = 0; tap < 9; ++tap, ++k) {
                        const int slot = k % C::S;
                        mbar_wait(&empty[slot], (uint32_t)(((k / C::S) & 1) ^ 1));
                        if (rank == 0) mbar_expect_tx(&full[slot], (uint32_t)(2 * C::HPX * C::ROWB));
                        tma2_load_im2col_4d(ring + slot * C::SLOT, &map_a, mapa_u32(&full[slot], 0), 0, 2 * w0, 2 * h, img,
                                            (uint16_t)(tap % 3), (uint16_t)(tap / 3));
                    }
                }
            } else
            for (int t = t_begin; t < t_end;) {
                const int col = t / Ho;
                const int h0 = t - col * Ho;
                const int len = min(t_end - t, Ho - h0);
                const int ip = col / P.tiles_x;
                const int w0 = (col - ip * P.tiles_x) * 128;
                const int img = 2 * ip + (int)rank;
                for (int i = 0; i < len + 2; ++i, ++k) {
                    const int slot = k % C::S;
                    mbar_wait(&empty[slot], (uint32_t)(((k / C::S) & 1) ^ 1));
                    if (rank == 0) mbar_expect_tx(&full[slot], (uint32_t)(2 * C::HPX * C::ROWB));
                    tma2_load_4d(ring + slot * C::SLOT, &map_a, mapa_u32(&full[slot], 0), 0, w0 - 1, h0 - 1 + i, img);
                    if (P.has_res && i >= 2) {
                        const int p = it & 1;
                        const uint32_t use = (uint32_t)(it >> 1);
                        mbar_wait(&stg_empty[p], (use & 1u) ^ 1u);
                        mbar_expect_tx(&res_full[p], (uint32_t)C::STG_BYTES);
                        for (int ch = 0; ch < C::NCHUNK; ++ch)
                            tma_load_4d(stg_base + p * C::STG_BYTES + ch * C::CHUNK_BYTES, &map_res, &res_full[p], ch * 64, w0,
                                        h0 + i - 2, img);
                        ++it;
                    }
                }
                if (!P.has_res) it += len;
                t += len;
            }
            // tail: multicast commits from the leader may still arrive on this CTA's empty barriers
            for (int j = 0; j < C::S; ++j, ++k) mbar_wait(&empty[k % C::S], (uint32_t)(((k / C::S) & 1) ^ 1));
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer: leader CTA, one thread
        if (rank == 0 && lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(256, COUT);
            mbar_wait(w_full, 0);
            tc_fence_after();
            const uint32_t w_addr = smem_u32(w_base);
            const uint32_t ring_addr = smem_u32(ring);
            int k0 = 0;
            int it = 0;
            if (STRIDE == 2) {
                for (int t = t_begin; t < t_end; ++t, ++it) {
                    const int p = it & 1;
                    const uint32_t use = (uint32_t)(it >> 1);
                    mbar_wait(&tmem_empty[p], (use & 1u) ^ 1u);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(p * COUT);
#pragma unroll 1
                    for (int tap = 0; tap < 9; ++tap, ++k0) {
                        const int slot = k0 % C::S;
                        mbar_wait(&full[slot], (uint32_t)((k0 / C::S) & 1));
                        tc_fence_after();
                        const uint64_t adesc = make_smem_desc(ring_addr + (uint32_t)(slot * C::SLOT), C::SBO, C::SWZ);
                        const uint64_t bdesc = make_smem_desc(w_addr + (uint32_t)(tap * C::WTAP), C::SBO, C::SWZ);
#pragma unroll
                        for (int kc = 0; kc < CIN / 16; ++kc)
                            umma2_bf16(d_tmem, adesc + (uint64_t)(kc * 2), bdesc + (uint64_t)(kc * 2), idesc, (uint32_t)((tap | kc) != 0));
                        umma2_commit_mc(&empty[slot], 3);
                    }
                    umma2_commit_mc(&tmem_full[p], 3);
                }
            } else
            for (int t = t_begin; t < t_end;) {
                const int col = t / Ho;
                const int h0 = t - col * Ho;
                const int len = min(t_end - t, Ho - h0);
                for (int r = 0; r < len; ++r, ++it) {
                    const int p = it & 1;
                    const uint32_t use = (uint32_t)(it >> 1);
                    mbar_wait(&tmem_empty[p], (use & 1u) ^ 1u);
                    for (int dy = (r == 0 ? 0 : 2); dy < 3; ++dy) {
                        const int kk = k0 + r + dy;
                        mbar_wait(&full[kk % C::S], (uint32_t)((kk / C::S) & 1));
                    }
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(p * COUT);
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy) {
                        const uint32_t row_addr = ring_addr + (uint32_t)(((k0 + r + dy) % C::S) * C::SLOT);
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            const uint64_t adesc = make_smem_desc(row_addr + (uint32_t)(dx * C::ROWB), C::SBO, C::SWZ);
                            const uint64_t bdesc = make_smem_desc(w_addr + (uint32_t)((dy * 3 + dx) * C::WTAP), C::SBO, C::SWZ);
#pragma unroll
                            for (int kc = 0; kc < CIN / 16; ++kc)
                                umma2_bf16(d_tmem, adesc + (uint64_t)(kc * 2), bdesc + (uint64_t)(kc * 2), idesc,
                                           (uint32_t)((dy | dx | kc) != 0));
                        }
                    }
                    umma2_commit_mc(&tmem_full[p], 3);
                    umma2_commit_mc(&empty[(k0 + r) % C::S], 3);          // input row r of the run is done
                    if (r == len - 1) {
                        umma2_commit_mc(&empty[(k0 + r + 1) % C::S], 3);
                        umma2_commit_mc(&empty[(k0 + r + 2) % C::S], 3);
                    }
                }
                k0 += len + 2;
                t += len;
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue: warps 2..9 of each CTA
        const int ew = warp - 2;
        const int q = warp & 3;                    // TMEM lane quadrant
        const int half = ew >> 2;                  // which half of the column groups
        const int row = q * 32 + lane;
        int it = 0;
        for (int t = t_begin; t < t_end; ++t, ++it) {
            const int col = t / Ho;
            const int h = t - col * Ho;
            const int ip = col / P.tiles_x;
            const int w0 = (col - ip * P.tiles_x) * 128;
            const int img = 2 * ip + (int)rank;
            const int p = it & 1;
            const uint32_t use = (uint32_t)(it >> 1);
            mbar_wait(&tmem_full[p], use & 1u);
            tc_fence_after();
            if (P.has_res) mbar_wait(&res_full[p], use & 1u);
            unsigned char* stg = stg_base + p * C::STG_BYTES;
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(p * COUT);
#pragma unroll 1
            for (int g = half; g < COUT / 32; g += 2) {
                uint32_t v[32];
                tmem_ld_32x32(t_addr + (uint32_t)(g * 32), v);
                tmem_ld_wait();
                const int c0 = g * 32;
                float y[32];
                const float4* pb = reinterpret_cast<const float4*>(s_par + c0);
                const float4* ps = reinterpret_cast<const float4*>(s_par + COUT + c0);
                const float4* pt = reinterpret_cast<const float4*>(s_par + 2 * COUT + c0);
#pragma unroll
                for (int k4 = 0; k4 < 8; ++k4) {
                    const float4 b4 = pb[k4], s4 = ps[k4], t4 = pt[k4];
                    const float2 za = add2_f32(make_float2(__uint_as_float(v[4 * k4 + 0]), __uint_as_float(v[4 * k4 + 1])), make_float2(b4.x, b4.y));
                    const float2 zb = add2_f32(make_float2(__uint_as_float(v[4 * k4 + 2]), __uint_as_float(v[4 * k4 + 3])), make_float2(b4.z, b4.w));
                    float2 la = mul2_f32(za, make_float2(0.2f, 0.2f)), lb = mul2_f32(zb, make_float2(0.2f, 0.2f));
                    la.x = fmaxf(la.x, za.x); la.y = fmaxf(la.y, za.y);        // leaky(z) = max(z, 0.2 z)
                    lb.x = fmaxf(lb.x, zb.x); lb.y = fmaxf(lb.y, zb.y);
                    const float2 ya = fma2_f32(la, make_float2(s4.x, s4.y), make_float2(t4.x, t4.y));
                    const float2 yb = fma2_f32(lb, make_float2(s4.z, s4.w), make_float2(t4.z, t4.w));
                    y[4 * k4 + 0] = ya.x; y[4 * k4 + 1] = ya.y; y[4 * k4 + 2] = yb.x; y[4 * k4 + 3] = yb.y;
                }
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
#pragma unroll
                for (int pc = 0; pc < 4; ++pc) {
                    uint4* dst = reinterpret_cast<uint4*>(rowp + (((piece0 + pc) ^ sw) << 4));
                    float* yy = y + pc * 8;
                    uint4 o;
                    o.x = pack2h(yy[0], yy[1]); o.y = pack2h(yy[2], yy[3]);
                    o.z = pack2h(yy[4], yy[5]); o.w = pack2h(yy[6], yy[7]);
                    *dst = o;
                }
            }
            // accumulator drained: release it to the leader's MMA thread (remote arrive from rank 1)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(&tmem_empty[p], 0));
            fence_proxy_async_smem();
            // drain the previous tile's store (other staging buffer) before anyone may pass the barrier
            if (warp == 2 && lane == 0 && !P.has_res) tma_store_wait_read();
            named_bar_sync(1, 256);
            if (warp == 2 && lane == 0) {
                if (img < P.n_img)
                    for (int ch = 0; ch < C::NCHUNK; ++ch)
                        tma_store_4d(&map_out, stg + ch * C::CHUNK_BYTES, ch * 64, w0, h, img);
                tma_store_commit();
                if (P.has_res) {
                    tma_store_wait_read();
                    mbar_arrive(&stg_empty[p]);
                }
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

template <int CIN, int COUT, int STRIDE>
static void launch_halo_t(y3_context* ctx, const ConvLaunch& L) {
    using C = HaloCfg<CIN, COUT, STRIDE>;
    static bool attr[64] = {};
    if (!attr[ctx->device & 63]) {
        Y3_CUDA(cudaFuncSetAttribute(k_conv_halo<CIN, COUT, STRIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        attr[ctx->device & 63] = true;
    }
    const ConvArgs& A = L.args;
    const long long total = (long long)((A.n_img + 1) / 2) * A.tiles_x * A.Ho;
    const int pairs = (int)std::min<long long>(total, (long long)(ctx->sm_count / 2) * C::CTAS_PER_SM);
    launch_pdl(k_conv_halo<CIN, COUT, STRIDE>, 2 * pairs, HALO_THREADS, C::SMEM, ctx->stream, L.map_a, L.map_b, L.map_out, L.map_res, L.args);
    Y3_LAUNCHED(ctx);
}

bool halo_supported(int cin, int cout_pad) { return (cin == 32 && cout_pad == 64) || (cin == 64 && cout_pad == 128); }

void launch_conv_halo(y3_context* ctx, const ConvLaunch& L) {
    const int s = L.args.stride;
    if (L.args.cin == 32 && L.bn == 64 && s == 1) launch_halo_t<32, 64, 1>(ctx, L);
    else if (L.args.cin == 64 && L.bn == 128 && s == 1) launch_halo_t<64, 128, 1>(ctx, L);
    else if (L.args.cin == 32 && L.bn == 64 && s == 2) launch_halo_t<32, 64, 2>(ctx, L);
    else if (L.args.cin == 64 && L.bn == 128 && s == 2) launch_halo_t<64, 128, 2>(ctx, L);
    else fail(Y3_ERR_UNSUPPORTED, "no halo conv kernel for Cin=%d Cout=%d stride %d", L.args.cin, L.bn, s);
}

}  // namespace y3
