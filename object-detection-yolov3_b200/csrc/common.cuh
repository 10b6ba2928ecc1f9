// common.cuh - context, error handling and small device helpers shared by all translation units.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include <map>
#include <memory>
#include <utility>
#include <stdlib.h>

#include <nvtx3/nvToolsExt.h>

#include "../../include/yolo3_b200.h"

namespace y3 {

struct Error {
    y3_status code;
    std::string msg;
};

// Thrown inside the library, caught at the extern "C" boundary (api.cu) - never crosses the ABI.
[[noreturn]] void fail(y3_status code, const char* fmt, ...);

#define Y3_CUDA(expr)                                                                       \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess)                                                              \
            ::y3::fail(Y3_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                       __FILE__, __LINE__);                                                 \
    } while (0)

#define Y3_CHECK(cond, code, ...)                      \
    do {                                               \
        if (!(cond)) ::y3::fail((code), __VA_ARGS__);  \
    } while (0)

// Growable device buffer (never shrinks; freed with the context).
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    void reserve(size_t bytes);
    void release();
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
    ~DevBuf() { release(); }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
};

struct PinnedBuf {
    void* p = nullptr;
    size_t cap = 0;
    void reserve(size_t bytes);
    void release();
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
    ~PinnedBuf() { release(); }
    PinnedBuf() = default;
    PinnedBuf(const PinnedBuf&) = delete;
    PinnedBuf& operator=(const PinnedBuf&) = delete;
};

struct EventTimer {
    cudaEvent_t a = nullptr, b = nullptr;
    void init();
    void destroy();
    void start(cudaStream_t s) { cudaEventRecord(a, s); }
    void stop(cudaStream_t s) { cudaEventRecord(b, s); }
    float ms();  // synchronises on b
};

struct Net;        // net.cu
struct PostProc;   // nms.cu
struct Tiler;      // tiles.cu

}  // namespace y3

struct y3_context {
    y3_config cfg{};
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    std::string last_error;
    int64_t kernels_launched = 0;
    y3_timings timings{};
    y3::Net* net = nullptr;
    y3::PostProc* post = nullptr;
    y3::Tiler* tiler = nullptr;
    y3::DevBuf stage_in, stage_out, stage_aux;   // host<->device staging for API calls
    // stage timers: event pairs recorded on the stream, resolved once at the end of the API call
    struct PhaseRec { cudaEvent_t a, b; float* dst; };
    std::vector<PhaseRec> phase_log;
    std::vector<cudaEvent_t> event_pool;
    cudaStream_t copy_stream = nullptr;          // H2D of the image band, overlapped with compute
    cudaStream_t post_stream = nullptr;          // candidates / sort / NMS / stitch of batch k while conv of k+1 runs
    y3::PinnedBuf pin_small;
    void* comm = nullptr;                        // y3::Comm (comm.cu): NCCL communicator of the sharded path
};

namespace y3 {

inline void count_launch(y3_context* c, int n = 1) { c->kernels_launched += n; }

// Checks the launch and counts it.
#define Y3_LAUNCHED(ctx)                \
    do {                                \
        Y3_CUDA(cudaGetLastError());    \
        ::y3::count_launch((ctx));      \
    } while (0)

// ---- stage timers: event pairs recorded on the current stream, resolved once at the end of an API call
inline cudaEvent_t take_event(y3_context* c) {
    if (!c->event_pool.empty()) { cudaEvent_t e = c->event_pool.back(); c->event_pool.pop_back(); return e; }
    cudaEvent_t e;
    Y3_CUDA(cudaEventCreate(&e));
    return e;
}
// Times a stage with two events on ctx->stream (whatever stream that is when start / stop run); nothing blocks
// until flush_phases().
// NVTX range on the host timeline (header-only NVTX3: a no-op unless a tool such as Nsight Systems is attached)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

struct Phase {
    y3_context* c;
    cudaEvent_t a;
    float* dst;
    bool open = true, ranged = false;
    // name (optional): the stage also becomes an NVTX range covering its enqueue
    Phase(y3_context* ctx, float* d, const char* name = nullptr) : c(ctx), a(take_event(ctx)), dst(d) {
        if (name) { nvtxRangePushA(name); ranged = true; }
        cudaEventRecord(a, c->stream);
    }
    void stop() {
        if (!open) return;
        open = false;
        cudaEvent_t b = take_event(c);
        cudaEventRecord(b, c->stream);
        c->phase_log.push_back({a, b, dst});
        if (ranged) { nvtxRangePop(); ranged = false; }
    }
    ~Phase() {
        if (open) c->event_pool.push_back(a);
        if (ranged) nvtxRangePop();
    }
};
inline void flush_phases(y3_context* c) {
    cudaStreamSynchronize(c->stream);
    for (auto& r : c->phase_log) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) *r.dst += t;
        c->event_pool.push_back(r.a);
        c->event_pool.push_back(r.b);
    }
    c->phase_log.clear();
    cudaGetLastError();
}

// ---- programmatic dependent launch for chains of short kernels (the post-processing pipeline): the next kernel of the
// stream may be scheduled while the current one drains; every chained kernel starts with pdl_chain_sync(), which waits
// until the grids it depends on have completed (memory visible) and then lets ITS successor be scheduled.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_chain_sync() {
    asm volatile("griddepcontrol.wait;\n" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
}
#endif
template <typename... KArgs, typename... Args>
inline void launch_chained(y3_context* ctx, void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream, Args&&... args) {
    static const bool pdl = getenv("Y3_NO_POST_PDL") == nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    Y3_CUDA(cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...));
    count_launch(ctx);
}

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
static inline int ilog2_ceil(uint64_t v) { int b = 0; while ((1ull << b) < v) ++b; return b; }

// ---- fp32 helpers that reproduce NumPy semantics exactly (no contraction, NaN propagation) ----
__device__ __forceinline__ float np_max(float a, float b) { return (a >= b || a != a) ? a : b; }
__device__ __forceinline__ float np_min(float a, float b) { return (a <= b || a != a) ? a : b; }

// bbox_utils.compute_iou (bbox_utils.py:200-214) for one pair, bit-exact with NumPy fp32:
//   inter = max(yb - yt, 0) * max(xr - xl, 0); union = (area_a + area_b) - inter; iou = inter / union
__device__ __forceinline__ float iou_exact(const float4 a, const float area_a, const float4 b, const float area_b) {
    const float xl = np_max(a.x, b.x);
    const float yt = np_max(a.y, b.y);
    const float xr = np_min(a.z, b.z);
    const float yb = np_min(a.w, b.w);
    const float dh = np_max(__fsub_rn(yb, yt), 0.0f);
    const float dw = np_max(__fsub_rn(xr, xl), 0.0f);
    const float inter = __fmul_rn(dh, dw);
    const float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
    return __fdiv_rn(inter, uni);
}
// single_class_nms survivor test `iou <= thr` (bbox_utils.py:233) negated, bit-exact, without the
// division for the common non-overlapping pair: inter == 0 gives iou = 0/union, i.e. +-0 (survives
// iff 0 <= thr) unless union is 0 or NaN (0/0 -> NaN -> suppressed).
__device__ __forceinline__ bool suppresses_exact(const float4 a, const float area_a, const float4 b, const float area_b,
                                                 const float thr) {
    const float xl = np_max(a.x, b.x);
    const float yt = np_max(a.y, b.y);
    const float xr = np_min(a.z, b.z);
    const float yb = np_min(a.w, b.w);
    const float dh = np_max(__fsub_rn(yb, yt), 0.0f);
    const float dw = np_max(__fsub_rn(xr, xl), 0.0f);
    const float inter = __fmul_rn(dh, dw);
    const float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
    if (inter == 0.0f) return !((uni != 0.0f) && (uni == uni) && (0.0f <= thr));
    return !(__fdiv_rn(inter, uni) <= thr);
}
// The same decision without the NaN-propagating selects and - outside a narrow band around the threshold - without
// the division, for a pair whose coordinates and areas are all finite (box_is_plain) and 0 < thr < inf:
//   RN(inter / uni) > thr  is certain when  inter > RN(RN(thr * uni) * (1 + 2^-20)),
//   RN(inter / uni) <= thr is certain when  inter < RN(RN(thr * uni) * (1 - 2^-20))
// (three roundings of relative size 2^-24 against a margin of 2^-20); anything else takes the exact path.
__device__ __forceinline__ bool box_is_plain(const float4 b, const float area) {
    const float inf = __int_as_float(0x7f800000);
    return fabsf(b.x) < inf && fabsf(b.y) < inf && fabsf(b.z) < inf && fabsf(b.w) < inf && fabsf(area) < inf;
}
__device__ __forceinline__ bool suppresses_plain(const float4 a, const float area_a, const float4 b, const float area_b,
                                                 const float thr) {
    const float xl = fmaxf(a.x, b.x);
    const float yt = fmaxf(a.y, b.y);
    const float xr = fminf(a.z, b.z);
    const float yb = fminf(a.w, b.w);
    const float dh = fmaxf(__fsub_rn(yb, yt), 0.0f);
    const float dw = fmaxf(__fsub_rn(xr, xl), 0.0f);
    const float inter = __fmul_rn(dh, dw);
    const float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
    if (inter == 0.0f) return !((uni != 0.0f) && (uni == uni) && (0.0f <= thr));
    if (uni > 1e-30f && uni < 1e30f && inter > 1e-30f && inter < 1e30f) {
        const float t = __fmul_rn(thr, uni);
        if (inter > __fmul_rn(t, 1.00000095367431640625f)) return true;
        if (inter < __fmul_rn(t, 0.99999904632568359375f)) return false;
    }
    return !(__fdiv_rn(inter, uni) <= thr);
}
__device__ __forceinline__ float box_area_exact(const float4 b) {
    return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
}

}  // namespace y3
