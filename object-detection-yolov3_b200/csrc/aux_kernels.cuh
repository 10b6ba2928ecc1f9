// aux_kernels.cuh - prototypes of the non-GEMM kernels (see aux_kernels.cu).
#pragma once
#include "common.cuh"

namespace y3 {

struct DecodeArgs {
    const float* head[3];     // fp32 [B, gh*gw, pitch] per scale (stride 32, 16, 8)
    int gh[3], gw[3];
    int row_start[3];         // first output row of each scale
    float stride_h[3], stride_w[3];
    float anchor_w[Y3_MAX_ANCHORS], anchor_h[Y3_MAX_ANCHORS];
    int na, nc, pitch, n_total, batch;
    float clip_w, clip_h;     // > 0: corners clipped to [0, clip_w] x [0, clip_h] (inference.py:62-65, np.clip) - decode_box only
};

#ifdef __CUDACC__
// reorg_layer + convert_feature_map_to_inference_detections (model.py:122-212), one output element.
// fp32 in the reference's operand order, no FMA contraction:
//   cx = (sigmoid(tx) + j) * stride ; w = exp(tw) * anchor_w ; x0 = cx - w / 2 ; x1 = cx + w / 2
__device__ __forceinline__ float sigmoid_f(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

// Stored channel order of one head cell (pitch floats): the A objectness logits first, then per anchor tx ty tw th and the
// class logits:  [obj_0 .. obj_{A-1} | a=0: tx ty tw th cls_0 .. | a=1: ... ].  It is a permutation of the detection
// layer's output channels (reference order a*(5+NC)+k), applied to the weight rows at load time, so the convolution writes
// it for free - and the threshold pass, which looks at the objectness of every row first, finds the A logits of a cell in
// ONE 32-byte sector instead of A different 128-byte lines (K2: 87 -> 29 MB of line traffic).
__host__ __device__ __forceinline__ int head_pos(int na, int nc, int ref_channel) {
    const int a = ref_channel / (5 + nc), k = ref_channel - a * (5 + nc);
    return k == 4 ? a : na + a * (4 + nc) + (k < 4 ? k : k - 1);
}
// pointer to the box + class logits of output row `row` of image b (hp[0..3] = tx ty tw th, hp[4 + c] = class c), the
// address of its objectness logit, and its grid cell / anchor / scale
__device__ __forceinline__ const float* head_row(const DecodeArgs& D, int b, int row, int* s_out, int* cell_out, int* a_out,
                                                 const float** obj_out) {
    int s = 0;
    if (row >= D.row_start[1]) s = 1;
    if (row >= D.row_start[2]) s = 2;
    const int lr = row - D.row_start[s];
    const int cell = lr / D.na;
    const int a = lr - cell * D.na;
    *s_out = s; *cell_out = cell; *a_out = a;
    const float* base = D.head[s] + ((long long)b * D.gh[s] * D.gw[s] + cell) * D.pitch;
    *obj_out = base + a;
    return base + D.na + a * (4 + D.nc);
}
// k in 0..3 : x0, y0, x1, y1
__device__ __forceinline__ float decode_corner(const DecodeArgs& D, const float* hp, int s, int cell, int a, int k) {
    const int axis = k & 1;                         // 0: x, 1: y
    const float tc = __ldg(hp + axis);
    const float ts = __ldg(hp + 2 + axis);
    const int gi = cell / D.gw[s], gj = cell - gi * D.gw[s];
    const float off = axis == 0 ? (float)gj : (float)gi;
    // model.py:127,157 - the (h,w) stride pair multiplies the (x,y) pair as written
    const float stride = axis == 0 ? D.stride_h[s] : D.stride_w[s];
    const float c = __fmul_rn(__fadd_rn(sigmoid_f(tc), off), stride);
    const float wh = __fmul_rn(expf(ts), axis == 0 ? D.anchor_w[a] : D.anchor_h[a]);
    const float half = __fdiv_rn(wh, 2.0f);
    return (k < 2) ? __fsub_rn(c, half) : __fadd_rn(c, half);
}
// all four corners of one row at once (the same operations as four decode_corner calls, each evaluated once)
__device__ __forceinline__ float4 decode_box(const DecodeArgs& D, const float* hp, int s, int cell, int a) {
    const float tx = __ldg(hp + 0), ty = __ldg(hp + 1), tw = __ldg(hp + 2), th = __ldg(hp + 3);
    const int gi = cell / D.gw[s], gj = cell - gi * D.gw[s];
    const float cx = __fmul_rn(__fadd_rn(sigmoid_f(tx), (float)gj), D.stride_h[s]);
    const float cy = __fmul_rn(__fadd_rn(sigmoid_f(ty), (float)gi), D.stride_w[s]);
    const float hw = __fdiv_rn(__fmul_rn(expf(tw), D.anchor_w[a]), 2.0f);
    const float hh = __fdiv_rn(__fmul_rn(expf(th), D.anchor_h[a]), 2.0f);
    float4 b = make_float4(__fsub_rn(cx, hw), __fsub_rn(cy, hh), __fadd_rn(cx, hw), __fadd_rn(cy, hh));
    if (D.clip_w > 0.f) {                         // np.clip(v, 0, hi) = minimum(maximum(v, 0), hi), NaN propagates
        b.x = np_min(np_max(b.x, 0.f), D.clip_w); b.z = np_min(np_max(b.z, 0.f), D.clip_w);
        b.y = np_min(np_max(b.y, 0.f), D.clip_h); b.w = np_min(np_max(b.w, 0.f), D.clip_h);
    }
    return b;
}
#endif

void launch_stem(y3_context* ctx, const float* in, __nv_bfloat16* out, const float* w, const float* bias,
                 const float* scale, const float* shift, int B, int H, int W, int cin);
// tensor-core stem (stem_tc.cu); returns false when the channel count is not covered
bool launch_stem_tc(y3_context* ctx, const float* in, __nv_bfloat16* out, const float* w, const float* bias, const float* scale,
                    const float* shift, int B, int H, int W, int cin);
// f16: pack as fp16 instead of bf16 (layers whose input tensor is fp16 - the tail after the first upsample)
// det_na > 0: a detection layer - output rows stored in head_pos() order
void pack_conv_weight(y3_context* ctx, const float* k, __nv_bfloat16* out, int taps, int cin, int cout, int cout_pad, bool f16 = false,
                      int det_na = 0, int det_nc = 0);
void permute_det_bias(y3_context* ctx, const float* ref_bias, float* out, int na, int nc);
void pack_convt_weight(y3_context* ctx, const float* k, __nv_bfloat16* out, long long n, bool f16 = false);
void compose_up(y3_context* ctx, const float* wy, const float* kt, const float* by, const float* bt, int c_up, int c_x, int c_r,
                int cout, __nv_bfloat16* w_out, float* b_out, bool x_f16 = false, bool r_f16 = false);
void bn_fold(y3_context* ctx, const float* g, const float* b, const float* m, const float* v, float* s, float* t, int c);
void heads_to_nchw(y3_context* ctx, const float* in, float* out, int B, int HW, int C, int pitch, int na, int nc);
void slice_to_nchw(y3_context* ctx, const __nv_bfloat16* in, float* out, int B, int H, int W, int C, int pitch, int coff, bool f16 = false);
void launch_decode(y3_context* ctx, const DecodeArgs& D, float* out);

}  // namespace y3
