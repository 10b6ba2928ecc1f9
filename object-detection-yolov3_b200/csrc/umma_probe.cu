// umma_probe.cu - hardware probe (test hook, not on the product path): does a K-major SWIZZLE_128B UMMA
// shared-memory descriptor accept a start address that is shifted by whole 128-byte rows (not aligned
// to the 1024-byte swizzle atom), and which `base_offset` makes it read what TMA wrote?
// D[128 x 64] = A[rows s .. s+127][64] * I  for several row shifts s and both base_offset conventions.
#include "common.cuh"
#include "../../include/yolo3_b200_probe.h"
#include "conv_tc.cuh"
#include "ptx.cuh"

namespace y3 {
using namespace ptx;

struct ProbeArgs { int shifts[8]; int n_shift; };

__global__ void __launch_bounds__(128, 1)
k_umma_probe(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, ProbeArgs P,
             float* __restrict__ out /*[2][n_shift][128][64]*/) {
#if defined(__CUDA_ARCH_FEAT_SM100_ALL) || defined(__CUDA_ARCH_FEAT_SM101_ALL)
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* sA = smem;                    // 512 rows x 128 B = 64 KB
    unsigned char* sB = smem + 65536;            // 64 rows x 128 B = 8 KB
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536 + 8192);
    uint64_t* mma_bar = bar + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(mma_bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc<64>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, 65536 + 8192);
        tma_load_2d(sA, &map_a, bar, 0, 0);
        tma_load_2d(sA + 32768, &map_a, bar, 0, 256);
        tma_load_2d(sB, &map_b, bar, 0, 0);
    }
    mbar_wait(bar, 0);
    uint32_t mma_phase = 0;
    for (int variant = 0; variant < 2; ++variant) {
        for (int si = 0; si < P.n_shift; ++si) {
            if (threadIdx.x == 0) {
                tc_fence_after();
                const uint32_t a_addr = smem_u32(sA) + (uint32_t)P.shifts[si] * 128u;
                uint64_t adesc = make_smem_desc(a_addr, 1024, SWZ_128B);
                if (variant == 1) adesc |= (uint64_t)((a_addr >> 7) & 7u) << 49;       // base_offset
                const uint64_t bdesc = make_smem_desc(smem_u32(sB), 1024, SWZ_128B);
                constexpr uint32_t idesc = make_idesc_bf16(128, 64);
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (uint32_t)(k != 0));
                umma_commit(mma_bar);
            }
            mbar_wait(mma_bar, mma_phase);
            mma_phase ^= 1u;
            tc_fence_after();
            const int row = warp * 32 + lane;
            float* o = out + (((size_t)variant * P.n_shift + si) * 128 + row) * 64;
            for (int g = 0; g < 2; ++g) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(g * 32), v);
                tmem_ld_wait();
                for (int k = 0; k < 32; ++k) o[g * 32 + k] = __uint_as_float(v[k]);
            }
            tc_fence_before();
            __syncthreads();
        }
    }
    if (warp == 0) tmem_dealloc<64>(tmem_base);
#endif
}

}  // namespace y3

extern "C" y3_status y3_debug_umma_rowshift(y3_handle h, const uint16_t* a_bf16 /*[512][64]*/, const int32_t* shifts, int32_t n_shift,
                                            float* out /*[2][n_shift][128][64]*/) {
    using namespace y3;
    if (!h || !a_bf16 || !shifts || !out || n_shift < 1 || n_shift > 8) return Y3_ERR_INVALID;
    try {
        Y3_CUDA(cudaSetDevice(h->device));
        DevBuf dA, dB, dO;
        dA.reserve(512 * 64 * 2); dB.reserve(64 * 64 * 2); dO.reserve((size_t)2 * n_shift * 128 * 64 * 4);
        std::vector<uint16_t> ident(64 * 64, 0);
        for (int i = 0; i < 64; ++i) ident[i * 64 + i] = 0x3f80;                 // bf16 1.0
        Y3_CUDA(cudaMemcpy(dA.p, a_bf16, 512 * 64 * 2, cudaMemcpyHostToDevice));
        Y3_CUDA(cudaMemcpy(dB.p, ident.data(), 64 * 64 * 2, cudaMemcpyHostToDevice));
        CUtensorMap ma, mb;
        { uint64_t d[2] = {64, 512}; uint64_t s[1] = {128}; uint32_t b[2] = {64, 256}; encode_tmap_bf16(&ma, dA.p, 2, d, s, b, 128); }
        { uint64_t d[2] = {64, 64}; uint64_t s[1] = {128}; uint32_t b[2] = {64, 64}; encode_tmap_bf16(&mb, dB.p, 2, d, s, b, 128); }
        ProbeArgs P{};
        P.n_shift = n_shift;
        for (int i = 0; i < n_shift; ++i) P.shifts[i] = shifts[i];
        const int smem = 1024 + 65536 + 8192 + 64;
        Y3_CUDA(cudaFuncSetAttribute(k_umma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        k_umma_probe<<<1, 128, smem, h->stream>>>(ma, mb, P, dO.as<float>());
        Y3_CUDA(cudaGetLastError());
        Y3_CUDA(cudaStreamSynchronize(h->stream));
        Y3_CUDA(cudaMemcpy(out, dO.p, (size_t)2 * n_shift * 128 * 64 * 4, cudaMemcpyDeviceToHost));
        return Y3_OK;
    } catch (const Error& e) {
        h->last_error = e.msg;
        cudaGetLastError();
        return e.code;
    }
}

// ------------------------------------------------------------------------------------------------
// Probe 2: im2col-mode TMA.  One load of 128 output pixels x 64 channels for tap (ow, oh) starting at
// input coordinate (w, h, n); the raw (swizzled) 16 KB tile is copied out for comparison on the host.
namespace y3 {
struct Im2colProbe { int c, w, h, n, ow, oh; };
__global__ void __launch_bounds__(128, 1)
k_im2col_probe(const __grid_constant__ CUtensorMap map, Im2colProbe P, uint4* __restrict__ out /*[1024] 16-B pieces*/) {
#if defined(__CUDA_ARCH_FEAT_SM100_ALL) || defined(__CUDA_ARCH_FEAT_SM101_ALL)
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384);
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x7fc07fc0u, 0x7fc07fc0u, 0x7fc07fc0u, 0x7fc07fc0u);
    if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); }
    ptx::fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
        ptx::mbar_expect_tx(bar, 16384);
        ptx::tma_load_im2col_4d(smem, &map, bar, P.c, P.w, P.h, P.n, (uint16_t)P.ow, (uint16_t)P.oh);
    }
    ptx::mbar_wait(bar, 0);
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[i] = reinterpret_cast<uint4*>(smem)[i];
#endif
}
}  // namespace y3

extern "C" y3_status y3_debug_im2col(y3_handle h, const uint16_t* x_bf16 /*NHWC*/, int32_t N, int32_t H, int32_t W, int32_t C,
                                     int32_t stride, int32_t pad_lo, int32_t pad_hi, int32_t ksize, const int32_t* probes /*[n][6]*/,
                                     int32_t n_probe, uint16_t* out /*[n][128][64] raw smem*/) {
    using namespace y3;
    if (!h || !x_bf16 || !probes || !out) return Y3_ERR_INVALID;
    try {
        Y3_CUDA(cudaSetDevice(h->device));
        DevBuf dX, dO;
        const size_t nx = (size_t)N * H * W * C;
        dX.reserve(nx * 2 + 131072); dO.reserve(16384);
        Y3_CUDA(cudaMemset(dX.p, 0, nx * 2 + 131072));
        Y3_CUDA(cudaMemcpy(dX.p, x_bf16, nx * 2, cudaMemcpyHostToDevice));
        CUtensorMap m;
        uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
        int lower[2] = {-pad_lo, -pad_lo};
        int upper[2] = {pad_hi - (ksize - 1), pad_hi - (ksize - 1)};
        encode_tmap_im2col_bf16(&m, dX.p, dims, str, lower, upper, 64, 128, (uint32_t)stride, 128);
        const int smem = 1024 + 16384 + 64;
        Y3_CUDA(cudaFuncSetAttribute(k_im2col_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        for (int i = 0; i < n_probe; ++i) {
            Im2colProbe P{probes[6 * i], probes[6 * i + 1], probes[6 * i + 2], probes[6 * i + 3], probes[6 * i + 4], probes[6 * i + 5]};
            k_im2col_probe<<<1, 128, smem, h->stream>>>(m, P, dO.as<uint4>());
            Y3_CUDA(cudaGetLastError());
            Y3_CUDA(cudaStreamSynchronize(h->stream));
            Y3_CUDA(cudaMemcpy(out + (size_t)i * 8192, dO.p, 16384, cudaMemcpyDeviceToHost));
        }
        return Y3_OK;
    } catch (const Error& e) {
        h->last_error = e.msg;
        cudaGetLastError();
        return e.code;
    }
}
