// ptx.cuh - thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05
// (alloc / mma / commit / ld) and the UMMA shared-memory + instruction descriptors.
// Bit layouts follow the PTX ISA "tcgen05" chapter (same fields CUTLASS' cute/arch/mma_sm100_desc.hpp
// documents): smem descriptor = start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48)
// | layout_type [61,64); instruction descriptor = c_format [4,6) | a_format [7,10) | b_format [10,13)
// | a_major 15 | b_major 16 | N>>3 [17,23) | M>>4 [24,29).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

namespace y3 { namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
#ifndef Y3_MBAR_SUSPEND_NS
#define Y3_MBAR_SUSPEND_NS 10000
#endif
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
#if Y3_MBAR_SUSPEND_NS > 0
    // suspend-time hint: the waiting thread sleeps in hardware until the phase completes (or the hint expires)
    // instead of re-issuing the poll - fewer wasted issue slots while the step is power-capped.  Also measured for the
    // single MMA issuer thread alone (round 2): a pure poll on the operand-ring barriers makes the long 3x3 layers
    // 3-5 % SLOWER (0.266 vs 0.259 ms at 128->256 @64^2) - the polls compete with the TMA's complete_tx updates.
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"((uint32_t)Y3_MBAR_SUSPEND_NS)
        : "memory");
#else
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
#endif
    return ok != 0;
}
// First 1024-byte aligned address of the dynamic shared memory.  Plain pointer arithmetic on the __shared__ array (not a
// round trip through uintptr_t): the compiler keeps the address space, so every access through the result is an
// LDS / STS instead of a generic LD / ST (which it also has to order against every other generic access).
__device__ __forceinline__ unsigned char* align_smem_1024(unsigned char* smem_raw) {
    return smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
}
// 16-byte shared-memory load the compiler cannot speculate: inside `if (has_res)` it stays inside (a plain LDS gets
// if-converted into an unconditional load + select, which puts an LDS round trip in front of every staging store of
// the layers WITHOUT a residual: +8 % on conv2d_2).
__device__ __forceinline__ uint4 lds128(const void* p) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(p)) : "memory");
    return v;
}
// Bounded wait: a broken pipeline traps (launch fails with an error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (Y3_MBAR_SUSPEND_NS > 0 ? (1u << 18) : (1u << 22))) {
            printf("y3: mbarrier timeout block (%d,%d,%d) thread %d bar %p parity %u\n", blockIdx.x, blockIdx.y,
                   blockIdx.z, threadIdx.x, (void*)bar, parity);
            __trap();
        }
    }
}

// ---------------------------------------------------------------- fences
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::
            "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
        "[%2];\n" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, "
        "%7}], [%2];\n" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];\n" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
// all but the most recent store group have finished READING shared memory
__device__ __forceinline__ void tma_store_wait_read_keep1() { asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory"); }

// ---------------------------------------------------------------- tcgen05 / TMEM
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(dst_smem)),
                 "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "n"(kCols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]; one thread issues for the CTA.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
                 : "memory");
}
// 32 lanes x 32 columns of fp32: thread t of the warp receives lane (base_lane + t), columns [col, col+32)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// ---------------------------------------------------------------- descriptors
enum : uint32_t { SWZ_128B = 2, SWZ_64B = 4 };   // UMMA layout_type encodings

// K-major operand tile in shared memory, rows of `row_bytes` (= swizzle span: 128 or 64), 8-row
// groups `8*row_bytes` apart (SBO).  `layout` is SWZ_128B or SWZ_64B.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffffu) >> 4);             // start address
    d |= (uint64_t)1 << 16;                                // LBO (ignored for swizzled K-major)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;     // SBO
    d |= (uint64_t)1 << 46;                                // descriptor version (Blackwell)
    d |= (uint64_t)layout << 61;
    return d;
}
// bf16 x bf16 -> fp32, both operands K-major, M x N tile
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t M, uint32_t N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// fp16 x fp16 -> fp32 (a_format = b_format = 0): same tensor-core rate, 11-bit significand (the fp16 tail, net.cu)
__host__ __device__ constexpr uint32_t make_idesc_f16(uint32_t M, uint32_t N) {
    return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// two floats -> packed 16-bit pair for an activation tensor: bf16 (round to nearest even) or fp16 with the
// magnitude clamped to the largest finite half (an overflow saturates instead of turning into inf)
__device__ __forceinline__ uint32_t pack_act2(float lo, float hi, int f16) {
    if (f16) {
        lo = fminf(fmaxf(lo, -65504.f), 65504.f);
        hi = fminf(fmaxf(hi, -65504.f), 65504.f);
        const __half2 v = __floats2half2_rn(lo, hi);
        return *reinterpret_cast<const uint32_t*>(&v);
    }
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
}

}}  // namespace y3::ptx

// ======================================================================= 2-CTA (cta_group::2) forms
namespace y3 { namespace ptx {

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// shared::cluster address of `p` (a local shared variable) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
// Remote (or local) arrive through a shared::cluster address.  Default semantics (release at CTA scope):
// the cluster-scope form (.release.cluster) costs a full cluster-wide memory fence per call (ncu: 22 % of the
// stall samples of the 1x1 layers were stall_membar on it); the TMEM hand-off it is used for is ordered by
// tcgen05.fence::before_thread_sync / after_thread_sync around the barrier, exactly as in CUTLASS.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_addr) : "memory");
}
// TMA loads issued by either CTA of a pair; complete_tx lands on the barrier at cluster address `bar`
__device__ __forceinline__ void tma2_load_2d(void* dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
        "[%2];\n" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma2_load_4d(void* dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
        "%6}], [%2];\n" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma2_load_5d(void* dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3,
                                             int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
        "%6, %7}], [%2];\n" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(dst_smem)), "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "n"(kCols) : "memory");
}
// D[tmem, both CTAs] (+)= A[smem of each CTA: its 128 rows] * B[smem: N/2 rows in each CTA]; leader thread only
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive (once) on the barrier at the same smem offset in every CTA of `mask` when the MMAs retire
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(
                     smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}

}}  // namespace y3::ptx

// ======================================================================= im2col-mode TMA loads
namespace y3 { namespace ptx {
// 128 consecutive OUTPUT pixels (flattened n,h,w; traversal stride and padding live in the tensor map) x one
// K chunk of channels, for filter tap (ow, oh).  (c, w, h, n) = channel chunk and the INPUT coordinate of the
// first pixel's window origin (w = wo*stride + lower_corner_w ...).
__device__ __forceinline__ void tma_load_im2col_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c, int w, int h, int n,
                                                   uint16_t ow, uint16_t oh) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
        "[%2], {%7, %8};\n" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n), "h"(ow), "h"(oh)
        : "memory");
}
__device__ __forceinline__ void tma2_load_im2col_4d(void* dst, const CUtensorMap* m, uint32_t bar, int c, int w, int h, int n,
                                                    uint16_t ow, uint16_t oh) {
    asm volatile(
        "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
        "%5, %6}], [%2], {%7, %8};\n" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(ow), "h"(oh)
        : "memory");
}
}}  // namespace y3::ptx

// ======================================================================= packed fp32 (FADD2 / FMUL2 / FFMA2)
namespace y3 { namespace ptx {
__device__ __forceinline__ float2 add2_f32(float2 a, float2 b) {
    float2 r;
    asm("{.reg .b64 ra, rb, rr; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.f32x2 rr, ra, rb; mov.b64 {%0,%1}, rr;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 mul2_f32(float2 a, float2 b) {
    float2 r;
    asm("{.reg .b64 ra, rb, rr; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mul.f32x2 rr, ra, rb; mov.b64 {%0,%1}, rr;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 fma2_f32(float2 a, float2 b, float2 c) {
    float2 r;
    asm("{.reg .b64 ra, rb, rc, rr; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; fma.rn.f32x2 rr, ra, rb, rc; "
        "mov.b64 {%0,%1}, rr;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}
}}  // namespace y3::ptx

// ======================================================================= programmatic dependent launch
namespace y3 { namespace ptx {
// Blocks until the grids this grid depends on have completed and their memory is visible (no-op when the
// kernel was launched without the programmatic-serialization attribute).
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
// Lets the next grid of the stream start its prologue on SMs this grid no longer occupies.
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }
}}  // namespace y3::ptx
