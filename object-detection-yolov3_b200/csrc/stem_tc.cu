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

template <int CIN>
__global__ void __launch_bounds__(128, (CIN == 1) ? 8 : 6)
k_stem_tc(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, const float* __restrict__ w /*[9*CIN][32]*/,
          const float* __restrict__ bias, const float* __restrict__ scale, const float* __restrict__ shift, int B, int H, int W) {
#if defined(__CUDA_ARCH_FEAT_SM100_ALL) || defined(__CUDA_ARCH_FEAT_SM101_ALL)
    constexpr int KREAL = 9 * CIN;
    constexpr int KP = KREAL <= 16 ? 16 : 32;                // padded K
    constexpr int KCH = KP / 8;                              // 16-byte chunks per row
    constexpr uint32_t SBO = KCH * 128;                      // bytes between 8-row groups
    __shared__ __align__(128) unsigned char sA[128 * KP * 2];
    __shared__ __align__(128) unsigned char sB[32 * KP * 2];
    __shared__ float4 s_p[3 * 8];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    // weights: B[n][k] = w[k][n], same interleaved layout (rows = output channels)
    for (int i = tid; i < 32 * KCH; i += 128) {
        const int n = i / KCH, kc = i - n * KCH;
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k0 = kc * 8 + 2 * j;
            const float a = k0 < KREAL ? w[k0 * 32 + n] : 0.f;
            const float b = k0 + 1 < KREAL ? w[(k0 + 1) * 32 + n] : 0.f;
            __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
            pk[j] = *reinterpret_cast<uint32_t*>(&v);
        }
        *reinterpret_cast<uint4*>(sB + (n >> 3) * SBO + kc * 128 + (n & 7) * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
    if (tid < 8) {
        s_p[tid] = reinterpret_cast<const float4*>(bias)[tid];
        s_p[8 + tid] = reinterpret_cast<const float4*>(scale)[tid];
        s_p[16 + tid] = reinterpret_cast<const float4*>(shift)[tid];
    }
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc<32>(&tmem_slot);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const long long npix = (long long)B * H * W;
    const long long n_tiles = (npix + 127) / 128;
    uint32_t phase = 0;
    // taps of this thread's pixel of tile `tile` (zeros outside the image = SAME padding)
    auto gather = [&](long long tile, float (&v)[KP]) {
#pragma unroll
        for (int k = 0; k < KP; ++k) v[k] = 0.f;
        const long long pix = tile * 128 + tid;
        if (tile < n_tiles && pix < npix) {
            const int x = (int)(pix % W);
            const long long t = pix / W;
            const int y = (int)(t % H);
            const int b = (int)(t / H);
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci) {
                const float* plane = in + ((long long)b * CIN + ci) * H * W;
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
                    const int yy = y + kh - 1;
                    const bool yok = (yy >= 0) && (yy < H);
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const int xx = x + kw - 1;
                        if (yok && xx >= 0 && xx < W) v[(kh * 3 + kw) * CIN + ci] = __ldg(plane + (long long)yy * W + xx);
                    }
                }
            }
        }
    };
    float v[KP], vn[KP];
    gather(blockIdx.x, vn);
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long pix = tile * 128 + tid;
        const bool ok = pix < npix;
#pragma unroll
        for (int k = 0; k < KP; ++k) v[k] = vn[k];
        unsigned char* rowp = sA + (tid >> 3) * SBO + (tid & 7) * 16;
#pragma unroll
        for (int kc = 0; kc < KCH; ++kc) {
            uint32_t pk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                __nv_bfloat162 h2 = __floats2bfloat162_rn(v[kc * 8 + 2 * j], v[kc * 8 + 2 * j + 1]);
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
        gather(tile + gridDim.x, vn);          // next tile's loads fly while this tile's MMA + epilogue run
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
                const float4 bb = s_p[c4], ss = s_p[8 + c4], tt = s_p[16 + c4];
                float z0 = __uint_as_float(acc[4 * c4 + 0]) + bb.x, z1 = __uint_as_float(acc[4 * c4 + 1]) + bb.y;
                float z2 = __uint_as_float(acc[4 * c4 + 2]) + bb.z, z3 = __uint_as_float(acc[4 * c4 + 3]) + bb.w;
                z0 = (z0 > 0.f ? z0 : 0.2f * z0) * ss.x + tt.x;
                z1 = (z1 > 0.f ? z1 : 0.2f * z1) * ss.y + tt.y;
                z2 = (z2 > 0.f ? z2 : 0.2f * z2) * ss.z + tt.z;
                z3 = (z3 > 0.f ? z3 : 0.2f * z3) * ss.w + tt.w;
                __nv_bfloat162 a = __floats2bfloat162_rn(z0, z1), c = __floats2bfloat162_rn(z2, z3);
                o[2 * c4] = *reinterpret_cast<uint32_t*>(&a);
                o[2 * c4 + 1] = *reinterpret_cast<uint32_t*>(&c);
            }
            uint4* dst = reinterpret_cast<uint4*>(out + pix * 32);
            dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
            dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
            dst[2] = make_uint4(o[8], o[9], o[10], o[11]);
            dst[3] = make_uint4(o[12], o[13], o[14], o[15]);
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
    const long long n_tiles = ((long long)B * H * W + 127) / 128;
    const int blocks = (int)std::min<long long>(n_tiles, (long long)ctx->sm_count * (cin == 1 ? 8 : 6));
    if (cin == 1) k_stem_tc<1><<<blocks, 128, 0, ctx->stream>>>(in, out, w, bias, scale, shift, B, H, W);
    else k_stem_tc<3><<<blocks, 128, 0, ctx->stream>>>(in, out, w, bias, scale, shift, B, H, W);
    Y3_LAUNCHED(ctx);
    return true;
}

}  // namespace y3
