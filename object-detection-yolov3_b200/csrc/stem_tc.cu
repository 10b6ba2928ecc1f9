// stem_tc.cu - first conv_layer (model.py:385: 3x3 s1 SAME, Cimg -> 32, bias -> leaky(0.2) -> BN) on the
// tensor cores.  K = 9*Cimg (9 or 27) is padded to 16 / 32.  A pixel of the NCHW fp32 input is 2-6 bytes, far
// below TMA's 16-byte granularity, so the CTA's own threads build the im2col rows: thread r gathers the 9*Cimg
// taps of pixel r (coalesced along x), converts to bf16 and writes them in the canonical NO-SWIZZLE K-major
// UMMA layout (8 rows x 16 B core matrices: byte = (r/8)*SBO + (k/8)*128 + (r%8)*16 + (k%8)*2).  One thread
// issues tcgen05.mma (M=128, N=32, K=16 per step), the four warps drain TMEM, apply the fused epilogue and
// store 64 contiguous bytes per pixel (NHWC bf16).  Small smem/TMEM footprint => several CTAs per SM overlap
// gather, MMA and epilogue of different tiles.
#include "aux_kernels.cuh"
#include "ptx.cuh"

namespace y3 {
using namespace ptx;

__device__ __forceinline__ uint64_t make_smem_desc_noswz(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;      // K-direction core-matrix stride
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;      // M/N-direction 8-row group stride
    d |= (uint64_t)1 << 46;                                  // descriptor version (Blackwell)
    return d;                                                // layout_type 0 = no swizzle (interleaved)
}

__device__ __forceinline__ void st_global_256(void* p, const uint32_t* v) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]),
                 "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}

// The bias rides in the GEMM: K columns KREAL / KREAL+1 of every A row hold 1.0 and the matching weight columns hold
// bf16(bias) and bf16(bias - bf16(bias)) (error <= 2^-17 |bias|), so the epilogue is leaky -> scale/shift only.
template <int CIN>
__global__ void __launch_bounds__(128, (CIN == 1) ? 8 : 6)
k_stem_tc(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, const float* __restrict__ w /*[9*CIN][32]*/,
          const float* __restrict__ bias, const float* __restrict__ scale, const float* __restrict__ shift, int B, int H, int W) {
#if defined(__CUDA_ARCH_FEAT_SM100_ALL) || defined(__CUDA_ARCH_FEAT_SM101_ALL)
    constexpr int KREAL = 9 * CIN;
    constexpr int KP = KREAL + 2 <= 16 ? 16 : 32;            // padded K (taps + 2 bias columns)
    constexpr int KCH = KP / 8;                              // 16-byte chunks per row
    constexpr uint32_t SBO = KCH * 128;                      // bytes between 8-row groups
    static_assert(KREAL + 2 <= KP, "no room for the bias columns");
    __shared__ __align__(128) unsigned char sA[128 * KP * 2];
    __shared__ __align__(128) unsigned char sB[32 * KP * 2];
    __shared__ float4 s_p[2 * 8];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    // weights: B[n][k] = w[k][n], same interleaved layout (rows = output channels)
    for (int i = tid; i < 32 * KCH; i += 128) {
        const int n = i / KCH, kc = i - n * KCH;
        const float bn = bias[n];
        const float bhi = __bfloat162float(__float2bfloat16_rn(bn));
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float ab[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int k = kc * 8 + 2 * j + e;
                ab[e] = k < KREAL ? w[k * 32 + n] : (k == KREAL ? bhi : (k == KREAL + 1 ? bn - bhi : 0.f));
            }
            __nv_bfloat162 v = __floats2bfloat162_rn(ab[0], ab[1]);
            pk[j] = *reinterpret_cast<uint32_t*>(&v);
        }
        *reinterpret_cast<uint4*>(sB + (n >> 3) * SBO + kc * 128 + (n & 7) * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
    if (tid < 8) {
        s_p[tid] = reinterpret_cast<const float4*>(scale)[tid];
        s_p[8 + tid] = reinterpret_cast<const float4*>(shift)[tid];
    }
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc<32>(&tmem_slot);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const uint32_t uW = (uint32_t)W, uH = (uint32_t)H;
    const uint32_t n_rows = (uint32_t)B * uH;                 // image rows of the whole batch
    const uint32_t npix = n_rows * uW;
    const uint32_t n_tiles = (npix + 127u) / 128u;
    // position of the tile's first pixel, advanced without divisions: (image b0, row y0, column x0)
    const uint32_t step = 128u * gridDim.x;
    const uint32_t step_rows = step / uW, step_x = step - step_rows * uW;
    const uint32_t step_b = step_rows / uH, step_y = step_rows - step_b * uH;
    uint32_t x0, y0, b0;
    {
        const uint32_t p0 = 128u * blockIdx.x;
        const uint32_t r0 = p0 / uW;
        x0 = p0 - r0 * uW; b0 = r0 / uH; y0 = r0 - b0 * uH;
    }
    // taps of this thread's pixel of the tile at (b0, y0, x0) (zeros outside the image = SAME padding)
    auto gather = [&](bool live, float (&v)[9 * CIN]) {
        uint32_t x = x0 + (uint32_t)tid, y = y0, b = b0;
        while (x >= uW) { x -= uW; if (++y == uH) { y = 0; ++b; } }
        live = live && (b < (uint32_t)B);
        const bool xl = x > 0, xr = x + 1 < uW;
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
            const float* pc = in + ((size_t)(b * CIN + ci) * uH + y) * uW + x;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const bool yok = live && (kh == 0 ? y > 0 : (kh == 2 ? y + 1 < uH : true));
                const float* pr = pc + (kh - 1) * W;
                v[(kh * 3 + 0) * CIN + ci] = (yok && xl) ? __ldg(pr - 1) : 0.f;
                v[(kh * 3 + 1) * CIN + ci] = yok ? __ldg(pr) : 0.f;
                v[(kh * 3 + 2) * CIN + ci] = (yok && xr) ? __ldg(pr + 1) : 0.f;
            }
        }
    };
    auto advance = [&]() {
        x0 += step_x; y0 += step_y; b0 += step_b;
        if (x0 >= uW) { x0 -= uW; ++y0; }
        if (y0 >= uH) { y0 -= uH; ++b0; }
    };
    float vn[9 * CIN];
    gather(blockIdx.x < n_tiles, vn);
    uint32_t phase = 0;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t pix = tile * 128u + (uint32_t)tid;
        const bool ok = pix < npix;
        unsigned char* rowp = sA + (tid >> 3) * SBO + (tid & 7) * 16;
#pragma unroll
        for (int kc = 0; kc < KCH; ++kc) {
            uint32_t pk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = kc * 8 + 2 * j;
                const float a = k < KREAL ? vn[k < KREAL ? k : 0] : ((k == KREAL || k == KREAL + 1) ? 1.f : 0.f);
                const float c = k + 1 < KREAL ? vn[k + 1 < KREAL ? k + 1 : 0] : ((k + 1 == KREAL || k + 1 == KREAL + 1) ? 1.f : 0.f);
                __nv_bfloat162 h2 = __floats2bfloat162_rn(a, c);
                pk[j] = *reinterpret_cast<uint32_t*>(&h2);
            }
            *reinterpret_cast<uint4*>(rowp + kc * 128) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            constexpr uint32_t idesc = make_idesc_bf16(128, 32);
            const uint64_t adesc = make_smem_desc_noswz(smem_u32(sA), 128, SBO);
            const uint64_t bdesc = make_smem_desc_noswz(smem_u32(sB), 128, SBO);
#pragma unroll
            for (int ks = 0; ks < KP / 16; ++ks)       // each K step consumes two core matrices = 256 bytes
                umma_bf16(tmem_base, adesc + (uint64_t)(ks * 16), bdesc + (uint64_t)(ks * 16), idesc, (uint32_t)(ks != 0));
            umma_commit(&bar);
        }
        advance();
        gather(tile + gridDim.x < n_tiles, vn);  // next tile's loads fly while this tile's MMA + epilogue run
        mbar_wait(&bar, phase);
        phase ^= 1u;
        tc_fence_after();
        uint32_t acc[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16), acc);
        tmem_ld_wait();
        if (ok) {
            uint32_t o[16];
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
                const float4 ss = s_p[c4], tt = s_p[8 + c4];
                const float2 za = make_float2(__uint_as_float(acc[4 * c4 + 0]), __uint_as_float(acc[4 * c4 + 1]));
                const float2 zb = make_float2(__uint_as_float(acc[4 * c4 + 2]), __uint_as_float(acc[4 * c4 + 3]));
                float2 la = mul2_f32(za, make_float2(0.2f, 0.2f)), lb = mul2_f32(zb, make_float2(0.2f, 0.2f));
                la.x = fmaxf(la.x, za.x); la.y = fmaxf(la.y, za.y);        // leaky(z) = max(z, 0.2 z)
                lb.x = fmaxf(lb.x, zb.x); lb.y = fmaxf(lb.y, zb.y);
                const float2 ya = fma2_f32(la, make_float2(ss.x, ss.y), make_float2(tt.x, tt.y));
                const float2 yb = fma2_f32(lb, make_float2(ss.z, ss.w), make_float2(tt.z, tt.w));
                __nv_bfloat162 a = __floats2bfloat162_rn(ya.x, ya.y), c = __floats2bfloat162_rn(yb.x, yb.y);
                o[2 * c4] = *reinterpret_cast<uint32_t*>(&a);
                o[2 * c4 + 1] = *reinterpret_cast<uint32_t*>(&c);
            }
            __nv_bfloat16* dst = out + (size_t)pix * 32;
            st_global_256(dst, o);
            st_global_256(dst + 16, o + 8);
        }
        tc_fence_before();
        __syncthreads();          // TMEM drained and sA free before the next tile
    }
    if (warp == 0) tmem_dealloc<32>(tmem_base);
#endif
}

bool launch_stem_tc(y3_context* ctx, const float* in, __nv_bfloat16* out, const float* w, const float* bias, const float* scale,
                    const float* shift, int B, int H, int W, int cin) {
    if (cin != 1 && cin != 3) return false;                  // other channel counts use the FP32-pipe stem
    if ((long long)B * H * W >= (1ll << 31)) return false;   // 32-bit pixel indices
    const long long n_tiles = ((long long)B * H * W + 127) / 128;
    const int blocks = (int)std::min<long long>(n_tiles, (long long)ctx->sm_count * (cin == 1 ? 8 : 6));
    if (cin == 1) k_stem_tc<1><<<blocks, 128, 0, ctx->stream>>>(in, out, w, bias, scale, shift, B, H, W);
    else k_stem_tc<3><<<blocks, 128, 0, ctx->stream>>>(in, out, w, bias, scale, shift, B, H, W);
    Y3_LAUNCHED(ctx);
    return true;
}

}  // namespace y3
