// aux_kernels.cu - the non-GEMM kernels around the conv stack:
//   k_stem            first conv_layer (model.py:385): 3x3 s1, Cimg (1 or 3) -> 32, read straight from
//                     the NCHW fp32 batch, fused bias -> leaky(0.2) -> BN, NHWC bf16 out.  K = 9*Cimg is
//                     9..27: far too thin for a 128-wide UMMA tile and the layer is output-bandwidth
//                     bound (64 B written per pixel for <= 27 reads), so it runs on the FP32 pipe.
//   k_pack_conv_w     Keras Conv2D kernel [kh,kw,Cin,Cout] fp32 -> [Cout_pad][kh*kw*Cin] bf16 (K-major)
//   k_pack_convt_w    Keras Conv2DTranspose kernel [2,2,Cout,Cin] fp32 -> 4 x [Cout][Cin] bf16
//   k_bn_fold         (gamma,beta,mean,var) -> s = gamma/sqrt(var+1e-3), t = beta - mean*s  (SURVEY Q1,Q3)
//   k_heads_to_nchw   [B,HW,pitch] fp32 -> NCHW fp32 (the parity point of y3_forward_heads)
//   k_decode          reorg_layer + convert_feature_map_to_inference_detections (model.py:122-212)
#include "aux_kernels.cuh"
#include <cuda_fp16.h>
#include <stdlib.h>

namespace y3 {

// ------------------------------------------------------------------------------------------ stem
template <int CIN>
__global__ void __launch_bounds__(128)
k_stem(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, const float* __restrict__ w /*[9*CIN][32]*/,
       const float* __restrict__ bias, const float* __restrict__ scale, const float* __restrict__ shift,
       int B, int H, int W) {
    // each thread produces TWO horizontally adjacent pixels x 32 channels: every weight float4 read from
    // shared memory feeds 8 FMAs, and the 3x4 input patch is shared by both pixels
    __shared__ float4 s_w[9 * CIN * 8];
    __shared__ float4 s_p[3 * 8];
    for (int i = threadIdx.x; i < 9 * CIN * 8; i += blockDim.x) s_w[i] = reinterpret_cast<const float4*>(w)[i];
    if (threadIdx.x < 8) {
        s_p[threadIdx.x] = reinterpret_cast<const float4*>(bias)[threadIdx.x];
        s_p[8 + threadIdx.x] = reinterpret_cast<const float4*>(scale)[threadIdx.x];
        s_p[16 + threadIdx.x] = reinterpret_cast<const float4*>(shift)[threadIdx.x];
    }
    __syncthreads();
    const int W2 = W >> 1;                                   // W is a multiple of 32
    const long long npair = (long long)B * H * W2;
    for (long long pp = (long long)blockIdx.x * blockDim.x + threadIdx.x; pp < npair;
         pp += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(pp % W2) * 2;
        const long long t = pp / W2;
        const int y = (int)(t % H);
        const int b = (int)(t / H);
        float acc0[32], acc1[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) { acc0[c] = 0.f; acc1[c] = 0.f; }
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
            const float* plane = in + ((long long)b * CIN + ci) * H * W;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const int yy = y + kh - 1;
                const bool yok = (yy >= 0) && (yy < H);
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int xx = x + j - 1;
                    v[j] = (yok && xx >= 0 && xx < W) ? __ldg(plane + (long long)yy * W + xx) : 0.f;
                }
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const float4* wr = s_w + ((kh * 3 + kw) * CIN + ci) * 8;
                    const float a0 = v[kw], a1 = v[kw + 1];
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4) {
                        const float4 ww = wr[c4];
                        acc0[4 * c4 + 0] = fmaf(a0, ww.x, acc0[4 * c4 + 0]); acc1[4 * c4 + 0] = fmaf(a1, ww.x, acc1[4 * c4 + 0]);
                        acc0[4 * c4 + 1] = fmaf(a0, ww.y, acc0[4 * c4 + 1]); acc1[4 * c4 + 1] = fmaf(a1, ww.y, acc1[4 * c4 + 1]);
                        acc0[4 * c4 + 2] = fmaf(a0, ww.z, acc0[4 * c4 + 2]); acc1[4 * c4 + 2] = fmaf(a1, ww.z, acc1[4 * c4 + 2]);
                        acc0[4 * c4 + 3] = fmaf(a0, ww.w, acc0[4 * c4 + 3]); acc1[4 * c4 + 3] = fmaf(a1, ww.w, acc1[4 * c4 + 3]);
                    }
                }
            }
        }
        const long long pix = ((long long)b * H + y) * W + x;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const float* acc = half ? acc1 : acc0;
            uint32_t o[16];
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
                const float4 bb = s_p[c4], ss = s_p[8 + c4], tt = s_p[16 + c4];
                float z0 = acc[4 * c4 + 0] + bb.x, z1 = acc[4 * c4 + 1] + bb.y;
                float z2 = acc[4 * c4 + 2] + bb.z, z3 = acc[4 * c4 + 3] + bb.w;
                z0 = (z0 > 0.f ? z0 : 0.2f * z0) * ss.x + tt.x;
                z1 = (z1 > 0.f ? z1 : 0.2f * z1) * ss.y + tt.y;
                z2 = (z2 > 0.f ? z2 : 0.2f * z2) * ss.z + tt.z;
                z3 = (z3 > 0.f ? z3 : 0.2f * z3) * ss.w + tt.w;
                __nv_bfloat162 a = __floats2bfloat162_rn(z0, z1), c = __floats2bfloat162_rn(z2, z3);
                o[2 * c4] = *reinterpret_cast<uint32_t*>(&a);
                o[2 * c4 + 1] = *reinterpret_cast<uint32_t*>(&c);
            }
            uint4* dst = reinterpret_cast<uint4*>(out + (pix + half) * 32);
            dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
            dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
            dst[2] = make_uint4(o[8], o[9], o[10], o[11]);
            dst[3] = make_uint4(o[12], o[13], o[14], o[15]);
        }
    }
}

void launch_stem(y3_context* ctx, const float* in, __nv_bfloat16* out, const float* w, const float* bias,
                 const float* scale, const float* shift, int B, int H, int W, int cin) {
    static const bool use_tc = getenv("Y3_STEM_FP32") == nullptr;
    if (use_tc && launch_stem_tc(ctx, in, out, w, bias, scale, shift, B, H, W, cin)) return;
    const long long npix = (long long)B * H * W / 2;        // two pixels per thread
    const int blocks = (int)std::min<long long>((npix + 127) / 128, (long long)ctx->sm_count * 64);
    if (cin == 1) k_stem<1><<<blocks, 128, 0, ctx->stream>>>(in, out, w, bias, scale, shift, B, H, W);
    else if (cin == 3) k_stem<3><<<blocks, 128, 0, ctx->stream>>>(in, out, w, bias, scale, shift, B, H, W);
    else if (cin == 2) k_stem<2><<<blocks, 128, 0, ctx->stream>>>(in, out, w, bias, scale, shift, B, H, W);
    else if (cin == 4) k_stem<4><<<blocks, 128, 0, ctx->stream>>>(in, out, w, bias, scale, shift, B, H, W);
    else fail(Y3_ERR_UNSUPPORTED, "stem supports 1..4 image channels, got %d", cin);
    Y3_LAUNCHED(ctx);
}

// ------------------------------------------------------------------------------------------ packing
// 16-bit operand element: bf16, or fp16 for layers of the fp16 tail (values beyond the half range saturate)
__device__ __forceinline__ __nv_bfloat16 to_operand16(float v, bool f16) {
    if (!f16) return __float2bfloat16_rn(v);
    const __half h = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
    return *reinterpret_cast<const __nv_bfloat16*>(&h);
}

__global__ void k_pack_conv_w(const float* __restrict__ k /*[taps,Cin,Cout]*/, __nv_bfloat16* __restrict__ out,
                              int taps, int cin, int cout, int cout_pad, int f16, int det_na, int det_nc) {
    const long long n = (long long)cout_pad * taps * cin;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = to_operand16(0.f, f16 != 0);
    // second sweep: rows land at their stored position - the identity (same thread, same element) except for detection
    // layers, which run as ONE block so that this barrier orders the zero fill before the permuted writes
    __syncthreads();
    const long long m = (long long)cout * taps * cin;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % cin);
        const long long r = i / cin;
        const int tap = (int)(r % taps);
        const int co = (int)(r / taps);                                       // reference output channel
        const int row = det_na > 0 ? head_pos(det_na, det_nc, co) : co;       // stored row
        out[((long long)row * taps + tap) * cin + ci] = to_operand16(k[((long long)tap * cin + ci) * cout + co], f16 != 0);
    }
}
__global__ void k_permute_det_bias(const float* __restrict__ ref_bias, float* __restrict__ out, int na, int nc) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < na * (5 + nc)) out[head_pos(na, nc, c)] = ref_bias[c];
}
__global__ void k_pack_convt_w(const float* __restrict__ k /*[4,Cout,Cin]*/, __nv_bfloat16* __restrict__ out, long long n, int f16) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = to_operand16(k[i], f16 != 0);
}
__global__ void k_bn_fold(const float* __restrict__ g, const float* __restrict__ b, const float* __restrict__ m,
                          const float* __restrict__ v, float* __restrict__ s, float* __restrict__ t, int c) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c) return;
    const float sc = g[i] / sqrtf(v[i] + 1e-3f);
    s[i] = sc;
    t[i] = b[i] - m[i] * sc;
}

// Fused upsample: Conv2DTranspose(k2,s2) followed by concat [up, route] and a 1x1 conv is one linear map per
// output phase (i,j):  W'_ij[k, ci] = sum_co Wy[co, k] * Kt[i,j,co,ci]  (fp32, rounded to bf16 once),
// route columns copied, bias' = by + sum_co Wy[co,k] * bt[co].   Wy: Keras [2C, cout]; Kt: Keras [2,2,C_up,C_x].
__global__ void k_compose_up(const float* __restrict__ wy, const float* __restrict__ kt, const float* __restrict__ by,
                             const float* __restrict__ bt, int c_up, int c_x, int c_r, int cout,
                             __nv_bfloat16* __restrict__ w_out /*[4][cout][c_x+c_r]*/, float* __restrict__ b_out,
                             int x_f16, int r_f16) {
    const int K = c_x + c_r;
    const long long n = 4LL * cout * K;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int kk = (int)(i % K);
        const long long r = i / K;
        const int k = (int)(r % cout);
        const int ij = (int)(r / cout);
        float acc;
        if (kk < c_x) {
            acc = 0.f;
            for (int co = 0; co < c_up; ++co)
                acc = fmaf(wy[(long long)co * cout + k], kt[((long long)ij * c_up + co) * c_x + kk], acc);
        } else {
            acc = wy[(long long)(c_up + kk - c_x) * cout + k];
        }
        w_out[i] = to_operand16(acc, (kk < c_x ? x_f16 : r_f16) != 0);     // each K segment in its tensor's format
        if (ij == 0 && kk == 0) {
            float b = by[k];
            for (int co = 0; co < c_up; ++co) b = fmaf(wy[(long long)co * cout + k], bt[co], b);
            b_out[k] = b;
        }
    }
}
void compose_up(y3_context* ctx, const float* wy, const float* kt, const float* by, const float* bt, int c_up, int c_x, int c_r,
                int cout, __nv_bfloat16* w_out, float* b_out, bool x_f16, bool r_f16) {
    const long long n = 4LL * cout * (c_x + c_r);
    k_compose_up<<<(int)std::min<long long>((n + 255) / 256, 8192), 256, 0, ctx->stream>>>(wy, kt, by, bt, c_up, c_x, c_r, cout, w_out, b_out,
                                                                                           x_f16 ? 1 : 0, r_f16 ? 1 : 0);
    Y3_LAUNCHED(ctx);
}

void pack_conv_weight(y3_context* ctx, const float* k, __nv_bfloat16* out, int taps, int cin, int cout, int cout_pad, bool f16, int det_na,
                      int det_nc) {
    const long long n = (long long)cout_pad * taps * cin;
    // one block: the zero fill of the padded rows and the (possibly permuted) row writes must not race
    if (det_na > 0) k_pack_conv_w<<<1, 1024, 0, ctx->stream>>>(k, out, taps, cin, cout, cout_pad, f16 ? 1 : 0, det_na, det_nc);
    else k_pack_conv_w<<<(int)std::min<long long>((n + 255) / 256, 4096), 256, 0, ctx->stream>>>(k, out, taps, cin, cout, cout_pad, f16 ? 1 : 0, 0, 0);
    Y3_LAUNCHED(ctx);
}
void permute_det_bias(y3_context* ctx, const float* ref_bias, float* out, int na, int nc) {
    k_permute_det_bias<<<(na * (5 + nc) + 127) / 128, 128, 0, ctx->stream>>>(ref_bias, out, na, nc);
    Y3_LAUNCHED(ctx);
}
void pack_convt_weight(y3_context* ctx, const float* k, __nv_bfloat16* out, long long n, bool f16) {
    k_pack_convt_w<<<(int)std::min<long long>((n + 255) / 256, 4096), 256, 0, ctx->stream>>>(k, out, n, f16 ? 1 : 0);
    Y3_LAUNCHED(ctx);
}
void bn_fold(y3_context* ctx, const float* g, const float* b, const float* m, const float* v, float* s, float* t, int c) {
    k_bn_fold<<<(c + 127) / 128, 128, 0, ctx->stream>>>(g, b, m, v, s, t, c);
    Y3_LAUNCHED(ctx);
}

// ------------------------------------------------------------------------------------------ heads
__global__ void k_heads_to_nchw(const float* __restrict__ in, float* __restrict__ out, int B, int HW, int C, int pitch, int na, int nc) {
    const long long n = (long long)B * C * HW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(i % HW);
        const long long r = i / HW;
        const int c = (int)(r % C);
        const int b = (int)(r / C);
        out[i] = in[((long long)b * HW + p) * pitch + head_pos(na, nc, c)];      // back to the reference's channel order
    }
}
void heads_to_nchw(y3_context* ctx, const float* in, float* out, int B, int HW, int C, int pitch, int na, int nc) {
    const long long n = (long long)B * C * HW;
    k_heads_to_nchw<<<(int)std::min<long long>((n + 255) / 256, 8192), 256, 0, ctx->stream>>>(in, out, B, HW, C, pitch, na, nc);
    Y3_LAUNCHED(ctx);
}

__global__ void k_slice_to_nchw(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, int B, int H, int W, int C,
                                int pitch, int coff, int f16) {
    const long long n = (long long)B * C * H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W);
        long long r = i / W;
        const int y = (int)(r % H); r /= H;
        const int c = (int)(r % C);
        const int b = (int)(r / C);
        const __nv_bfloat16 e = in[(((long long)b * H + y) * W + x) * pitch + coff + c];
        out[i] = f16 ? __half2float(*reinterpret_cast<const __half*>(&e)) : __bfloat162float(e);
    }
}
void slice_to_nchw(y3_context* ctx, const __nv_bfloat16* in, float* out, int B, int H, int W, int C, int pitch, int coff, bool f16) {
    const long long n = (long long)B * C * H * W;
    k_slice_to_nchw<<<(int)std::min<long long>((n + 255) / 256, 8192), 256, 0, ctx->stream>>>(in, out, B, H, W, C, pitch, coff, f16 ? 1 : 0);
    Y3_LAUNCHED(ctx);
}

// ------------------------------------------------------------------------------------------ decode
// One thread per output element of [B, N, 5+NC].  A head stored NHWC with channel a*(5+NC)+k is already
// in output row order, so this is a pure element-wise map (helpers in aux_kernels.cuh).
__global__ void __launch_bounds__(256)
k_decode(DecodeArgs D, float* __restrict__ out) {
    const int E = 5 + D.nc;
    const long long per_img = (long long)D.n_total * E;
    const long long n = per_img * D.batch;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / per_img);
        const long long r = i - (long long)b * per_img;
        const int row = (int)(r / E);
        const int k = (int)(r - (long long)row * E);
        int s, cell, a;
        const float* ho;
        const float* hp = head_row(D, b, row, &s, &cell, &a, &ho);
        out[i] = (k == 4) ? sigmoid_f(__ldg(ho)) : (k > 4) ? sigmoid_f(__ldg(hp + k - 1)) : decode_corner(D, hp, s, cell, a, k);
    }
}
void launch_decode(y3_context* ctx, const DecodeArgs& D, float* out) {
    const long long n = (long long)D.n_total * (5 + D.nc) * D.batch;
    const int blocks = (int)std::min<long long>((n + 255) / 256, (long long)ctx->sm_count * 16);
    k_decode<<<blocks, 256, 0, ctx->stream>>>(D, out);
    Y3_LAUNCHED(ctx);
}

}  // namespace y3
