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
};

void launch_stem(y3_context* ctx, const float* in, __nv_bfloat16* out, const float* w, const float* bias,
                 const float* scale, const float* shift, int B, int H, int W, int cin);
void pack_conv_weight(y3_context* ctx, const float* k, __nv_bfloat16* out, int taps, int cin, int cout, int cout_pad);
void pack_convt_weight(y3_context* ctx, const float* k, __nv_bfloat16* out, long long n);
void bn_fold(y3_context* ctx, const float* g, const float* b, const float* m, const float* v, float* s, float* t, int c);
void heads_to_nchw(y3_context* ctx, const float* in, float* out, int B, int HW, int C, int pitch);
void slice_to_nchw(y3_context* ctx, const __nv_bfloat16* in, float* out, int B, int H, int W, int C, int pitch, int coff);
void launch_decode(y3_context* ctx, const DecodeArgs& D, float* out);

}  // namespace y3
