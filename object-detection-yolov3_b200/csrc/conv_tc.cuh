// conv_tc.cuh - host-side description of one tcgen05 implicit-GEMM convolution launch.
#pragma once
#include "ptx.cuh"
#include <cuda.h>
#include "common.cuh"
#include <stdlib.h>
#include <utility>

namespace y3 {

// Everything the kernel needs besides the four tensor maps.
struct ConvArgs {
    int tiles_x, tiles_y, tiles_per_img, n_img;
    int n_tiles_n, total_tiles;
    int BH, BW;              // spatial patch of one M tile (BH*BW <= 128 rows)
    int taps, kwn;           // 9/3 for 3x3, 1/1 for 1x1
    int cin, kchunks;        // channels per tap, cin / BK
    int stride, pad;         // 1|2 ; left/top zero padding (TF SAME: s1 k3 -> 1, s2 -> 0)
    int a_cpitch;            // channel pitch of the input buffer (phase offset of the stride-2 view)
    int k_split;             // K chunks [0,k_split) come from map_a, the rest from map_a2 (fused upsample+concat)
    int im2col;              // A operand through an im2col-mode tensor map: M tile = 128 consecutive output pixels
    int im_ho, im_wo;        //   output grid of one image
    int im_stride, im_lower; //   traversal stride and lower pixel-box corner (= -pad_before)
    int has_res, linear, out_f32;
    int in_f16, in2_f16;     // operand format of the K chunks read through map_a / map_a2 (and of their weights): fp16 instead of bf16
    int out_f16;             // the output activation tensor is fp16 (the fp16 tail after the first upsample, net.cu)
    int Ho, Wo;
    const float* bias;       // [cout_pad]
    const float* scale;      // [cout_pad]  gamma / sqrt(var + eps)       (unused when linear)
    const float* shift;      // [cout_pad]  beta - mean * scale
    float* out32;            // fp32 NHWC output (detection heads), pitch floats per pixel
    long long out32_pitch;
    int cout_valid;
};

struct ConvLaunch {
    CUtensorMap map_a, map_a2, map_b, map_out, map_res;
    ConvArgs args;
    int bn, bk;              // template selection (two_cta: bn = N of the pair UMMA, 128 or 256)
    int two_cta;             // use the cta_group::2 kernel (conv_tc2.cu)
    int halo;                // weights-stationary halo-row kernel (conv_halo.cu): 3x3 s1, Cin 32/64
    int grid;
};

// Encodes a tiled tensor map (bf16) through the driver entry point fetched at runtime.
void encode_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes /*rank-1*/, const uint32_t* box, int swizzle_bytes);

void launch_conv(y3_context* ctx, const ConvLaunch& L);

// Optional launch with programmatic stream serialization (Y3_PDL=1): the kernel may become resident while its
// predecessor in the stream drains; every thread executes griddepcontrol.wait before it touches activations.
// Measured on B200 (round 1): no gain for this stack (12.88 vs 12.83 ms per 128 tiles) - the persistent CTAs hold
// every SM until they exit, so there is nothing to overlap - hence off by default.
// Small batches (the captured batch-1 forward) turn it on per call: there a kernel is a handful of tiles, most SMs are
// free, and overlapping the next kernel's prologue (barrier init, TMEM allocation, tensor-map prefetch) with the tail
// of the current one is worth ~6 % of the K1 forward (1.33 -> 1.25 ms).
inline bool& pdl_for_small_batches() { static thread_local bool on = false; return on; }
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream, Args&&... args) {
    static const bool pdl_env = getenv("Y3_PDL") != nullptr;
    const bool pdl = pdl_env || pdl_for_small_batches();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    Y3_CUDA(cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...));
}
void launch_conv2(y3_context* ctx, const ConvLaunch& L);
bool launch_conv2h(y3_context* ctx, const ConvLaunch& L);   // half-staged variant (conv_tc2h.cu); false = not used
void launch_conv_halo(y3_context* ctx, const ConvLaunch& L);
// stem (1 image channel) + conv2d_1 fused (conv_stem1.cu); L = conv2d_1's halo-form launch
void launch_stem_conv1(y3_context* ctx, const ConvLaunch& L, const float* in, const float* stem_w_host, const float* stem_bias_host,
                       const float* stem_scale_host, const float* stem_shift_host, int H, int W);
bool halo_supported(int cin, int cout_pad);

// 32 epilogue values of one pixel -> four 16-byte pieces of its (swizzled) staging row: piece pc goes to
// rowp + (((piece0 + pc) ^ sw) << 4).  F16 = the tensor is stored as fp16 (saturating) instead of bf16; a template
// argument so that only one of the two conversions is emitted per group.
template <bool F16>
__device__ __forceinline__ void stage_row_32(const float (&y)[32], unsigned char* rowp, int piece0, int sw) {
#pragma unroll
    for (int pc = 0; pc < 4; ++pc) {
        const float* yy = y + pc * 8;
        uint4 o;
        o.x = ptx::pack_act2(yy[0], yy[1], F16); o.y = ptx::pack_act2(yy[2], yy[3], F16);
        o.z = ptx::pack_act2(yy[4], yy[5], F16); o.w = ptx::pack_act2(yy[6], yy[7], F16);
        *reinterpret_cast<uint4*>(rowp + (((piece0 + pc) ^ sw) << 4)) = o;
    }
}

// One 32-channel group of a conv epilogue: y = BN(LeakyReLU_0.2(acc + bias)), or acc + bias for a linear layer
// (reference model.py:29-39: Conv2D bias -> LeakyReLU(0.2) -> BatchNorm folded to scale/shift), in packed f32x2
// arithmetic.  LINEAR is a template argument so that the layer kind is one branch per group instead of a predicate
// on every element (the if-converted form cost ~7 instructions per value, this one ~2.5).
// b / s / t = the group's 32 bias / scale / shift values in global memory (16-byte aligned, read-only path).
template <bool LINEAR>
__device__ __forceinline__ void epilogue_math_32(const uint32_t (&v)[32], const float* __restrict__ b, const float* __restrict__ s,
                                                 const float* __restrict__ t, float (&y)[32]) {
    using namespace ptx;
#pragma unroll
    for (int k4 = 0; k4 < 8; ++k4) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(b) + k4);
        float2 za = add2_f32(make_float2(__uint_as_float(v[4 * k4 + 0]), __uint_as_float(v[4 * k4 + 1])), make_float2(b4.x, b4.y));
        float2 zb = add2_f32(make_float2(__uint_as_float(v[4 * k4 + 2]), __uint_as_float(v[4 * k4 + 3])), make_float2(b4.z, b4.w));
        if (!LINEAR) {
            const float4 s4 = __ldg(reinterpret_cast<const float4*>(s) + k4);
            const float4 t4 = __ldg(reinterpret_cast<const float4*>(t) + k4);
            float2 la = mul2_f32(za, make_float2(0.2f, 0.2f)), lb = mul2_f32(zb, make_float2(0.2f, 0.2f));
            la.x = fmaxf(la.x, za.x); la.y = fmaxf(la.y, za.y);        // leaky(z) = max(z, 0.2 z)
            lb.x = fmaxf(lb.x, zb.x); lb.y = fmaxf(lb.y, zb.y);
            za = fma2_f32(la, make_float2(s4.x, s4.y), make_float2(t4.x, t4.y));
            zb = fma2_f32(lb, make_float2(s4.z, s4.w), make_float2(t4.z, t4.w));
        }
        y[4 * k4 + 0] = za.x; y[4 * k4 + 1] = za.y; y[4 * k4 + 2] = zb.x; y[4 * k4 + 3] = zb.y;
    }
}

}  // namespace y3

namespace y3 {
// im2col-mode tensor map over an NHWC bf16 tensor: dims (C, W, H, N); lower/upper = pixel-box corners (W, H);
// pixels = output pixels per load (M tile), channels = K chunk; stride = traversal stride (W, H).
void encode_tmap_im2col_bf16(CUtensorMap* map, const void* base, const uint64_t* dims /*4*/, const uint64_t* strides_bytes /*3*/,
                             const int* lower /*2*/, const int* upper /*2*/, uint32_t channels, uint32_t pixels, uint32_t stride,
                             int swizzle_bytes);
}
