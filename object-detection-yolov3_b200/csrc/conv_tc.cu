// conv_tc.cu - the convolution of YoloV3.conv_layer / detection_layer / upsample_2x
// (reference model.py:29-39, 108-120, 94-105) as ONE persistent, warp-specialised sm_100a kernel:
//
//   implicit GEMM  D[M = pixels, N = Cout] = A[M, K = taps*Cin] * W[N, K]^T      bf16 x bf16 -> fp32
//
//   * activations are NHWC bf16.  An M tile is a BH x BW spatial patch of one image (BH*BW <= 128
//     rows); for filter tap (kh,kw) the A tile is that patch shifted by the tap, fetched with ONE
//     tiled TMA load whose out-of-bounds elements are zero-filled by the hardware - this is the
//     TF "SAME" padding: stride 1 pads (1,1); stride 2 pads (0 before, 1 after).
//     Stride-2 layers read the input through a 5-D view (pw*C+c, W/2, ph, H/2, N) so that the
//     even/odd phases are separate coordinates and no element stride is needed.
//     1x1 layers are launched with the "patch" 1 x 128 over the flattened pixel axis.
//   * weights are [Cout][taps*Cin] bf16 (K-major), fetched with a 2-D TMA box.
//   * warp 0 = TMA producer, warp 1 = tcgen05.mma issuer, warps 2-5 = epilogue.
//     smem ring of STAGES (A,B) tiles with full/empty mbarriers; the fp32 accumulator lives in
//     TMEM and is double-buffered (2 x BN columns) so the epilogue of tile i overlaps the MMAs of
//     tile i+1.  The CTA is persistent: tiles are strided over the grid, n-tiles fastest so that
//     concurrently running CTAs share A tiles in L2.
//   * epilogue (reference order):  z = acc + bias;  a = z > 0 ? z : 0.2 z;
//     y = a * s + t  with s = gamma/sqrt(var+1e-3), t = beta - mean*s;  y += X (block input, bf16
//     tile prefetched by TMA into the output staging buffer);  round to bf16 once; TMA store.
//     `linear` layers (detection heads, transposed-conv taps) do y = acc + bias; heads are stored
//     as fp32 straight from registers.
//   * route concat is zero-copy: the output tensor map addresses a channel slice of a wider
//     NHWC buffer; the transposed conv is 4 such GEMMs scattered by the output map's strides.
#include "conv_tc.cuh"
#include "ptx.cuh"
#include <algorithm>

namespace y3 {
using namespace ptx;

static constexpr int CONV_THREADS = 192;
static constexpr int TILE_M = 128;

template <int BN, int BK>
struct ConvCfg {
    static constexpr int A_BYTES = TILE_M * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int OC = BN < 64 ? BN : 64;        // channels per staging chunk (<= 128 B rows)
    static constexpr int OCB = OC * 2;
    static constexpr int NCHUNK = BN / OC;
    static constexpr int CHUNK_BYTES = TILE_M * OCB;
    static constexpr int STG_BYTES = TILE_M * BN * 2;
    // small configs run 2 CTAs per SM (their tiles are latency-chained, not bandwidth-bound per CTA)
    static constexpr int CTAS_PER_SM = (BN == 128) ? 1 : 2;
    static constexpr int STAGES = (BN == 128 && BK == 64) ? 4 : (BK == 64 ? 3 : 4);
    static constexpr int BAR_BYTES = (2 * STAGES + 8) * 8 + 16;
    static constexpr int SMEM = 1024 + STAGES * STAGE_BYTES + 2 * STG_BYTES + BAR_BYTES;
    static constexpr uint32_t TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
    static constexpr uint32_t SWZ = (BK == 64) ? (uint32_t)SWZ_128B : (uint32_t)SWZ_64B;
    static constexpr uint32_t SBO = 8 * BK * 2;
    static_assert(STAGE_BYTES % 1024 == 0 && A_BYTES % 1024 == 0, "operand tiles must stay 1024-B aligned");
    static_assert((TMEM_COLS & (TMEM_COLS - 1)) == 0 && TMEM_COLS <= 512, "TMEM columns: power of two <= 512");
    static_assert(SMEM * CTAS_PER_SM <= 232448 - 1024 * CTAS_PER_SM, "exceeds 227 KB of shared memory");
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

template <int BN, int BK>
__global__ void __launch_bounds__(CONV_THREADS, ConvCfg<BN, BK>::CTAS_PER_SM)
k_conv_tc(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a2,
          const __grid_constant__ CUtensorMap map_b,
          const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_res,
          const ConvArgs P) {
#if defined(__CUDA_ARCH_FEAT_SM100_ALL) || defined(__CUDA_ARCH_FEAT_SM101_ALL)
    using C = ConvCfg<BN, BK>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = align_smem_1024(smem_raw);
    unsigned char* stage_base = smem;
    unsigned char* stg_base = smem + C::STAGES * C::STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(stg_base + 2 * C::STG_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + C::STAGES;
    uint64_t* tmem_full = bars + 2 * C::STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* res_full = tmem_full + 4;
    uint64_t* stg_empty = tmem_full + 6;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 8);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

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
            mbar_init(&tmem_empty[p], 4);
            mbar_init(&res_full[p], 1);
            mbar_init(&stg_empty[p], 1);
        }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<C::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_launch_dependents();
    griddep_wait();              // activations of the previous layer are complete and visible from here on

    const int rows = P.BH * P.BW;
    const int k_iters = P.taps * P.kchunks;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x, ++it) {
                const int nt = t % P.n_tiles_n;
                const int mt = t / P.n_tiles_n;
                const int img = mt / P.tiles_per_img;
                const int r = mt - img * P.tiles_per_img;
                const int ty = r / P.tiles_x;
                const int x0 = (r - ty * P.tiles_x) * P.BW;
                const int y0 = ty * P.BH;
                const int n0 = nt * BN;
                int im_w = 0, im_h = 0, im_n = 0;
                if (P.im2col) {                               // x0 = first flattened output pixel of this tile
                    const int per = P.im_ho * P.im_wo;
                    im_n = x0 / per;
                    const int rem = x0 - im_n * per;
                    const int oh_ = rem / P.im_wo;
                    im_h = oh_ * P.im_stride + P.im_lower;
                    im_w = (rem - oh_ * P.im_wo) * P.im_stride + P.im_lower;
                }
                const int p = it & 1;
                const uint32_t use = (uint32_t)(it >> 1);
                if (P.has_res) {
                    mbar_wait(&stg_empty[p], (use & 1u) ^ 1u);
                    mbar_expect_tx(&res_full[p], (uint32_t)(rows * BN * 2));
                    for (int ch = 0; ch < C::NCHUNK; ++ch)
                        tma_load_4d(stg_base + p * C::STG_BYTES + ch * C::CHUNK_BYTES, &map_res, &res_full[p],
                                    n0 + ch * C::OC, x0, y0, img);
                }
                for (int tap = 0; tap < P.taps; ++tap) {
                    const int kh = tap / P.kwn;
                    const int kw = tap - kh * P.kwn;
                    for (int kc = 0; kc < P.kchunks; ++kc) {
                        mbar_wait(&empty[stage], phase ^ 1u);
                        unsigned char* sa = stage_base + stage * C::STAGE_BYTES;
                        mbar_expect_tx(&full[stage], (uint32_t)(rows * BK * 2 + C::B_BYTES));
                        if (P.im2col) {
                            tma_load_im2col_4d(sa, &map_a, &full[stage], kc * BK, im_w, im_h, im_n, (uint16_t)kw, (uint16_t)kh);
                        } else if (P.stride == 1) {
                            if (kc < P.k_split) tma_load_4d(sa, &map_a, &full[stage], kc * BK, x0 + kw - P.pad, y0 + kh - P.pad, img);
                            else tma_load_4d(sa, &map_a2, &full[stage], (kc - P.k_split) * BK, x0 + kw - P.pad, y0 + kh - P.pad, img);
                        }
                        else
                            tma_load_5d(sa, &map_a, &full[stage], (kw & 1) * P.a_cpitch + kc * BK, x0 + (kw >> 1), kh & 1,
                                        y0 + (kh >> 1), img);
                        tma_load_2d(sa + C::A_BYTES, &map_b, &full[stage], tap * P.cin + kc * BK, n0);
                        if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (one thread)
        if (lane == 0) {
            const uint32_t idesc1 = P.in_f16 ? make_idesc_f16(TILE_M, BN) : make_idesc_bf16(TILE_M, BN);
            const uint32_t idesc2 = P.in2_f16 ? make_idesc_f16(TILE_M, BN) : make_idesc_bf16(TILE_M, BN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x, ++it) {
                const int p = it & 1;
                const uint32_t use = (uint32_t)(it >> 1);
                mbar_wait(&tmem_empty[p], (use & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(p * BN);
                int kc = 0;
                for (int ki = 0; ki < k_iters; ++ki) {
                    const uint32_t idesc = kc < P.k_split ? idesc1 : idesc2;
                    if (++kc == P.kchunks) kc = 0;
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(stage_base + stage * C::STAGE_BYTES);
                    const uint64_t adesc = make_smem_desc(a_addr, C::SBO, C::SWZ);
                    const uint64_t bdesc = make_smem_desc(a_addr + C::A_BYTES, C::SBO, C::SWZ);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                                  (uint32_t)((ki | k) != 0));
                    umma_commit(&empty[stage]);                 // frees the smem slot when the MMAs retire
                    if (ki == k_iters - 1) umma_commit(&tmem_full[p]);
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue (warps 2..5)
        const int q = warp & 3;                    // TMEM lane quadrant this warp may access
        const int row = q * 32 + lane;
        const int by = row / P.BW;
        const int bx = row - by * P.BW;
        int it = 0;
        for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x, ++it) {
            const int nt = t % P.n_tiles_n;
            const int mt = t / P.n_tiles_n;
            const int img = mt / P.tiles_per_img;
            const int r = mt - img * P.tiles_per_img;
            const int ty = r / P.tiles_x;
            const int x0 = (r - ty * P.tiles_x) * P.BW;
            const int y0 = ty * P.BH;
            const int n0 = nt * BN;
            const int p = it & 1;
            const uint32_t use = (uint32_t)(it >> 1);
            mbar_wait(&tmem_full[p], use & 1u);
            tc_fence_after();
            if (P.has_res) mbar_wait(&res_full[p], use & 1u);
            unsigned char* stg = stg_base + p * C::STG_BYTES;
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(p * BN);
            const bool pix_ok = (row < rows) && (y0 + by < P.Ho) && (x0 + bx < P.Wo);
#pragma unroll 1
            for (int g = 0; g < BN / 32; ++g) {
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
                    // 32 channels = 4 pieces of 16 B inside staging chunk `ch`
                    const int cc = g * 32;
                    const int ch = cc / C::OC;
                    const int piece0 = (cc - ch * C::OC) >> 3;
                    unsigned char* rowp = stg + ch * C::CHUNK_BYTES + row * C::OCB;
                    const int sw = (C::OCB == 128) ? (row & 7) : ((row >> 1) & 3);
                    if (P.has_res) {                     // block input: all four pieces requested first, one round trip
#pragma unroll
                        for (int pc = 0; pc < 4; ++pc) {
                            const uint4 x = lds128(rowp + (((piece0 + pc) ^ sw) << 4));
                            float* yy = y + pc * 8;
                            yy[0] += bf16lo(x.x); yy[1] += bf16hi(x.x); yy[2] += bf16lo(x.y); yy[3] += bf16hi(x.y);
                            yy[4] += bf16lo(x.z); yy[5] += bf16hi(x.z); yy[6] += bf16lo(x.w); yy[7] += bf16hi(x.w);
                        }
                    }
                    if (P.out_f16) stage_row_32<true>(y, rowp, piece0, sw);
                    else stage_row_32<false>(y, rowp, piece0, sw);
                }
            }
            // accumulator drained: hand the TMEM buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[p]);
            if (!P.out_f32) {
                fence_proxy_async_smem();                  // generic-proxy smem writes -> visible to TMA
                // the store of the previous tile (other staging buffer) must have finished reading smem before
                // ANY thread passes this barrier and starts writing that buffer for the next tile
                if (warp == 2 && lane == 0 && !P.has_res) tma_store_wait_read();
                named_bar_sync(1, 128);
                if (warp == 2 && lane == 0) {
                    for (int ch = 0; ch < C::NCHUNK; ++ch)
                        tma_store_4d(&map_out, stg + ch * C::CHUNK_BYTES, n0 + ch * C::OC, x0, y0, img);
                    tma_store_commit();
                    if (P.has_res) {
                        tma_store_wait_read();             // residual prefetch may overwrite this buffer
                        mbar_arrive(&stg_empty[p]);
                    }                                      // (no residual: drained lazily before the next barrier,
                                                           //  so this store overlaps the next tile's epilogue)

                }
            }
        }
    }
    if (warp == 2 && lane == 0) tma_store_wait_read();     // smem must outlive the last TMA store's reads
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<C::TMEM_COLS>(tmem_base);
#else
    (void)P;
    __trap();   // this library is sm_100a only
#endif
}

// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled_t get_encode() {
    static PFN_encodeTiled_t fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        Y3_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        Y3_CHECK(p && q == cudaDriverEntryPointSuccess, Y3_ERR_CUDA, "cuTensorMapEncodeTiled not available");
        fn = reinterpret_cast<PFN_encodeTiled_t>(p);
    }
    return fn;
}

void encode_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bdim[5], estr[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bdim[i] = box[i]; estr[i] = 1; }
    for (int i = 0; i < rank - 1; ++i) gstr[i] = strides_bytes[i];
    const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64  ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle_bytes == 32  ? CU_TENSOR_MAP_SWIZZLE_32B
                                                       : CU_TENSOR_MAP_SWIZZLE_NONE;
    const CUresult r = get_encode()(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base),
                                    gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    Y3_CHECK(r == CUDA_SUCCESS, Y3_ERR_CUDA,
             "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu,%llu,%llu] box [%u,%u,%u,%u,%u] "
             "stride0 %llu swizzle %d base %p",
             (int)r, rank, (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 0),
             (unsigned long long)(rank > 2 ? gdim[2] : 0), (unsigned long long)(rank > 3 ? gdim[3] : 0),
             (unsigned long long)(rank > 4 ? gdim[4] : 0), bdim[0], rank > 1 ? bdim[1] : 0, rank > 2 ? bdim[2] : 0,
             rank > 3 ? bdim[3] : 0, rank > 4 ? bdim[4] : 0, (unsigned long long)(rank > 1 ? gstr[0] : 0),
             swizzle_bytes, base);
}


typedef CUresult (*PFN_encodeIm2col_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                       CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

void encode_tmap_im2col_bf16(CUtensorMap* map, const void* base, const uint64_t* dims, const uint64_t* strides_bytes,
                             const int* lower, const int* upper, uint32_t channels, uint32_t pixels, uint32_t stride,
                             int swizzle_bytes) {
    static PFN_encodeIm2col_t fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        Y3_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &q));
        Y3_CHECK(p && q == cudaDriverEntryPointSuccess, Y3_ERR_CUDA, "cuTensorMapEncodeIm2col not available");
        fn = reinterpret_cast<PFN_encodeIm2col_t>(p);
    }
    cuuint64_t gdim[4] = {dims[0], dims[1], dims[2], dims[3]};
    cuuint64_t gstr[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
    cuuint32_t estr[4] = {1, stride, stride, 1};
    int lo[2] = {lower[0], lower[1]}, up[2] = {upper[0], upper[1]};
    const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64  ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE;
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, lo, up, channels, pixels,
                          estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    int drv = 0;
    if (r == CUDA_SUCCESS && cudaDriverGetVersion(&drv) == cudaSuccess && drv <= 13010) {
        // same workaround CUTLASS applies (cute/atom/copy_traits_sm90_im2col.hpp) for tensors below 128 KB
        const uint64_t bytes = (dims[3] - 1) * strides_bytes[2] + (dims[2] - 1) * strides_bytes[1] + (dims[1] - 1) * strides_bytes[0] + dims[0] * 2;
        if (bytes < 131072) reinterpret_cast<uint64_t*>(map)[1] &= ~(1llu << 21);
    }
    Y3_CHECK(r == CUDA_SUCCESS, Y3_ERR_CUDA,
             "cuTensorMapEncodeIm2col failed (%d): dims [%llu,%llu,%llu,%llu] lower [%d,%d] upper [%d,%d] ch %u px %u stride %u",
             (int)r, (unsigned long long)gdim[0], (unsigned long long)gdim[1], (unsigned long long)gdim[2],
             (unsigned long long)gdim[3], lo[0], lo[1], up[0], up[1], channels, pixels, stride);
}

template <int BN, int BK>
static void launch_t(y3_context* ctx, const ConvLaunch& L) {
    using C = ConvCfg<BN, BK>;
    static bool attr[64] = {};
    if (!attr[ctx->device & 63]) {
        Y3_CUDA(cudaFuncSetAttribute(k_conv_tc<BN, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        attr[ctx->device & 63] = true;
    }
    const int grid = std::min(L.args.total_tiles, ctx->sm_count * C::CTAS_PER_SM);
    launch_pdl(k_conv_tc<BN, BK>, grid, CONV_THREADS, C::SMEM, ctx->stream, L.map_a, L.map_a2, L.map_b, L.map_out, L.map_res, L.args);
    Y3_LAUNCHED(ctx);
}

void launch_conv(y3_context* ctx, const ConvLaunch& L) {
    if (L.halo) { launch_conv_halo(ctx, L); return; }
    if (L.two_cta) { launch_conv2(ctx, L); return; }
    if (L.bn == 128 && L.bk == 64) launch_t<128, 64>(ctx, L);
    else if (L.bn == 64 && L.bk == 64) launch_t<64, 64>(ctx, L);
    else if (L.bn == 32 && L.bk == 64) launch_t<32, 64>(ctx, L);
    else if (L.bn == 64 && L.bk == 32) launch_t<64, 32>(ctx, L);
    else fail(Y3_ERR_UNSUPPORTED, "no conv kernel for BN=%d BK=%d", L.bn, L.bk);
}

}  // namespace y3
