// tiles.cu - tile slicing + per-tile normalisation (front-end) and seam stitching (back-end) of
// inference_tiled.py on the device.
//
// Front-end  (inference_tiled.py:29-100, 202-212; imagereader.py:34-46)
//   tile grid: zone = tile - 2r, tiles start every `zone` pixels, crop [i-r, i+zone+r) clamped to the
//   image and np.pad(mode='reflect') back to the tile size; r = 0 on an axis whose tile covers the
//   image.  The RECORDED origin is the clamped one (SURVEY Q12 - border tiles are shifted by +r; kept).
//   Each tile is z-scored with its own mean / population std over all channels (Q14); std <= 1 means
//   "subtract the mean only".  16 CTAs per tile: shifted fp64 sums (exact for integer images), then a
//   second kernel writes (x - mean) / std as NCHW fp32.  The image stays resident in
//   HBM in its source dtype; the reflect is index arithmetic, no padded copy is ever made.
// Back-end   (inference_tiled.py:235-301)
//   ghost-band ownership by box centre, origin add (fp32), np.round (half-to-even) -> int32,
//   centre-inside-image filter, clamp to [0, size-1]; ordered compaction keeps the reference's order.
#include "tiles.cuh"
#include <stdlib.h>

namespace y3 {

std::vector<TileGeo> plan_tiles(int64_t H, int64_t W, int th, int tw, int edge, int* ry_out, int* rx_out) {
    Y3_CHECK(th > 0 && tw > 0 && th % 32 == 0 && tw % 32 == 0, Y3_ERR_INVALID, "tile size %dx%d must be a multiple of 32", th, tw);
    Y3_CHECK(edge >= 0 && edge % 32 == 0, Y3_ERR_INVALID, "edge range %d must be a multiple of 32", edge);
    const int ry = th >= H ? 0 : edge, rx = tw >= W ? 0 : edge;
    const int zy = th - 2 * ry, zx = tw - 2 * rx;
    Y3_CHECK(zy > 0 && zx > 0, Y3_ERR_INVALID, "tile %dx%d leaves no zone of responsibility with edge range %d", th, tw, edge);
    std::vector<TileGeo> v;
    for (int64_t i = 0; i < H; i += zy)
        for (int64_t j = 0; j < W; j += zx) {
            TileGeo g;
            int64_t ys = i - ry, ye = i + zy + ry, xs = j - rx, xe = j + zx + rx;
            g.pre_y = ys < 0 ? (int)-ys : 0;
            g.pre_x = xs < 0 ? (int)-xs : 0;
            g.y0 = (int)std::max<int64_t>(ys, 0); g.x0 = (int)std::max<int64_t>(xs, 0);
            g.y1 = (int)std::min<int64_t>(ye, H); g.x1 = (int)std::min<int64_t>(xe, W);
            g.rec_x = g.x0; g.rec_y = g.y0;
            v.push_back(g);
        }
    if (ry_out) *ry_out = ry;
    if (rx_out) *rx_out = rx;
    return v;
}

// np.pad(mode='reflect') index: q relative to the crop start, n = crop length
__device__ __forceinline__ int reflect_idx(int q, int n) {
    if (n == 1) return 0;
    const int period = 2 * (n - 1);
    q %= period;
    if (q < 0) q += period;
    return q < n ? q : period - q;
}

template <typename T>
__device__ __forceinline__ float load_px(const void* img, long long idx) { return (float)reinterpret_cast<const T*>(img)[idx]; }

__device__ __forceinline__ float load_any(const void* img, int dtype, long long idx) {
    switch (dtype) {
        case Y3_U8: return load_px<uint8_t>(img, idx);
        case Y3_U16: return load_px<uint16_t>(img, idx);
        case Y3_I32: return load_px<int32_t>(img, idx);
        default: return load_px<float>(img, idx);
    }
}

__device__ __forceinline__ double block_sum(double v, double* s_red) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_red[wid] = v;
    __syncthreads();
    if (wid == 0) {
        double t = lane < (int)(blockDim.x >> 5) ? s_red[lane] : 0.0;
        for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) s_red[32] = t;
    }
    __syncthreads();
    return s_red[32];
}

static constexpr int TILE_SPLIT = 16;      // CTAs per tile (one CTA per tile left 116 of 148 SMs idle)

// source column of tile column tx (np.pad 'reflect' only at the borders; the interior is a plain offset)
__device__ __forceinline__ int tile_src(int t, int pre, int n, int origin) {
    const int q = t - pre;
    return origin + (((unsigned)q < (unsigned)n) ? q : reflect_idx(q, n));
}

// 1-channel uint16 tiles whose columns are a plain, 16-byte aligned span of the source rows (no reflection in x):
// both passes then move 8 pixels per load (the common case: every tile that does not touch the left / right border)
__device__ __forceinline__ bool tile_rows_vectorisable(const void* img, int dtype, int C, int W, int tw, const TileGeo& g, int nx) {
    return (reinterpret_cast<uintptr_t>(img) & 15) == 0 && dtype == Y3_U16 && C == 1 && g.pre_x == 0 && nx == tw && (tw & 7) == 0 && (W & 7) == 0 && (g.x0 & 7) == 0;
}

// pass 1: per-tile sum(x - s) and sum((x - s)^2) in fp64, s = the tile's first element (a shift that
// removes the cancellation of the one-pass variance; for integer images every partial sum is an exact
// integer < 2^53, so the result does not depend on the order of the atomics).
// Rows of the (c, ty) plane are strided over the CTAs of a tile, threads run along tx (coalesced).
__global__ void __launch_bounds__(512)
k_tile_stats(const void* __restrict__ img, int dtype, long long row_lo, int W, int C, const TileGeo* __restrict__ geo,
             int th, int tw, double* __restrict__ sums /*[count][2]*/) {
    __shared__ double s_red[33];
    const TileGeo g = geo[blockIdx.y];
    const int ny = g.y1 - g.y0, nx = g.x1 - g.x0;
    const double shift = (double)load_any(img, dtype, ((long long)(tile_src(0, g.pre_y, ny, g.y0) - row_lo) * W + tile_src(0, g.pre_x, nx, g.x0)) * C);
    double a1 = 0.0, a2 = 0.0;
    if (tile_rows_vectorisable(img, dtype, C, W, tw, g, nx)) {
        // 1-channel uint16 tile without horizontal reflection: 8 pixels per 16-byte load, exact integer sums
        const int cpr = tw >> 3;                                // 16-byte chunks per tile row
        const int ishift = (int)shift;
        long long s1 = 0;
        unsigned long long s2 = 0;
        for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < th * cpr; idx += gridDim.x * blockDim.x) {
            const int ty = idx / cpr, ch = idx - ty * cpr;
            const long long src = (long long)(tile_src(ty, g.pre_y, ny, g.y0) - row_lo) * W + g.x0 + 8 * ch;
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(img) + src));
            const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int d0 = (int)(w4[k] & 0xffffu) - ishift, d1 = (int)(w4[k] >> 16) - ishift;
                s1 += d0 + d1;
                s2 += (unsigned long long)((unsigned)d0 * (unsigned)d0) + (unsigned long long)((unsigned)d1 * (unsigned)d1);   // |d| < 2^16
            }
        }
        a1 = (double)s1;                                        // exact: |s1| < 2^53
        a2 = (double)s2;
    } else
    for (int r = blockIdx.x; r < C * th; r += gridDim.x) {
        const int c = r / th, ty = r - c * th;
        const long long rowbase = (long long)(tile_src(ty, g.pre_y, ny, g.y0) - row_lo) * W;
        for (int tx = threadIdx.x; tx < tw; tx += blockDim.x) {
            const double d = (double)load_any(img, dtype, (rowbase + tile_src(tx, g.pre_x, nx, g.x0)) * C + c) - shift;
            a1 += d;
            a2 += d * d;
        }
    }
    a1 = block_sum(a1, s_red);
    a2 = block_sum(a2, s_red);
    if (threadIdx.x == 0) {
        atomicAdd(&sums[2 * blockIdx.y], a1);
        atomicAdd(&sums[2 * blockIdx.y + 1], a2);
    }
}

// pass 2: (x - mean) / std, or x - mean when std <= 1 (imagereader.py:38-44), NCHW fp32
__global__ void __launch_bounds__(512)
k_tile_write(const void* __restrict__ img, int dtype, long long row_lo, int W, int C, const TileGeo* __restrict__ geo,
             int th, int tw, const double* __restrict__ sums, float* __restrict__ out, float* __restrict__ stats) {
    const TileGeo g = geo[blockIdx.y];
    const int ny = g.y1 - g.y0, nx = g.x1 - g.x0;
    const int n_el = th * tw * C;
    const double shift = (double)load_any(img, dtype, ((long long)(tile_src(0, g.pre_y, ny, g.y0) - row_lo) * W + tile_src(0, g.pre_x, nx, g.x0)) * C);
    const double m1 = sums[2 * blockIdx.y] / (double)n_el;
    double var = sums[2 * blockIdx.y + 1] / (double)n_el - m1 * m1;
    if (var < 0.0) var = 0.0;
    const float mu = (float)(shift + m1);
    const float sd = (float)sqrt(var);
    float* o = out + (long long)blockIdx.y * n_el;
    const bool center_only = sd <= 1.0f;
    if (tile_rows_vectorisable(img, dtype, C, W, tw, g, nx)) {
        const int cpr = tw >> 3;
        for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < th * cpr; idx += gridDim.x * blockDim.x) {
            const int ty = idx / cpr, ch = idx - ty * cpr;
            const long long src = (long long)(tile_src(ty, g.pre_y, ny, g.y0) - row_lo) * W + g.x0 + 8 * ch;
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(img) + src));
            const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
            float f[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float d0 = __fsub_rn((float)(w4[k] & 0xffffu), mu), d1 = __fsub_rn((float)(w4[k] >> 16), mu);
                f[2 * k] = center_only ? d0 : __fdiv_rn(d0, sd);
                f[2 * k + 1] = center_only ? d1 : __fdiv_rn(d1, sd);
            }
            float4* dst = reinterpret_cast<float4*>(o + (long long)ty * tw + 8 * ch);
            dst[0] = make_float4(f[0], f[1], f[2], f[3]);
            dst[1] = make_float4(f[4], f[5], f[6], f[7]);
        }
    } else
    for (int r = blockIdx.x; r < C * th; r += gridDim.x) {
        const int c = r / th, ty = r - c * th;
        const long long rowbase = (long long)(tile_src(ty, g.pre_y, ny, g.y0) - row_lo) * W;
        float* orow = o + (long long)r * tw;
        for (int tx = threadIdx.x; tx < tw; tx += blockDim.x) {
            const float d = __fsub_rn(load_any(img, dtype, (rowbase + tile_src(tx, g.pre_x, nx, g.x0)) * C + c), mu);
            orow[tx] = center_only ? d : __fdiv_rn(d, sd);
        }
    }
    if (stats && blockIdx.x == 0 && threadIdx.x == 0) { stats[2 * blockIdx.y] = mu; stats[2 * blockIdx.y + 1] = sd; }
}

void launch_tile_norm(y3_context* ctx, const void* img_dev, int dtype, long long row_lo, int W, int C,
                      const TileGeo* geo_dev, int count, int th, int tw, float* out, float* stats, double* sums_scratch) {
    if (count <= 0) return;
    Y3_CUDA(cudaMemsetAsync(sums_scratch, 0, (size_t)count * 16, ctx->stream));
    dim3 grid(TILE_SPLIT, count);
    const int threads = tw >= 512 ? 512 : (tw >= 256 ? 256 : 128);
    k_tile_stats<<<grid, threads, 0, ctx->stream>>>(img_dev, dtype, row_lo, W, C, geo_dev, th, tw, sums_scratch);
    Y3_LAUNCHED(ctx);
    k_tile_write<<<grid, threads, 0, ctx->stream>>>(img_dev, dtype, row_lo, W, C, geo_dev, th, tw, sums_scratch, out, stats);
    Y3_LAUNCHED(ctx);
}

// raw (un-normalised) tiles in the source dtype, HWC - the return value of convert_image_to_tiles
__global__ void __launch_bounds__(256)
k_tile_raw(const unsigned char* __restrict__ img, int esize, long long row_lo, int W, int C, const TileGeo* __restrict__ geo,
           int th, int tw, unsigned char* __restrict__ out) {
    const TileGeo g = geo[blockIdx.y];
    const int ny = g.y1 - g.y0, nx = g.x1 - g.x0;
    const int n_px = th * tw;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n_px; p += gridDim.x * blockDim.x) {
        const int ty = p / tw, tx = p - ty * tw;
        const int sy = g.y0 + reflect_idx(ty - g.pre_y, ny);
        const int sx = g.x0 + reflect_idx(tx - g.pre_x, nx);
        const unsigned char* src = img + (((long long)(sy - row_lo) * W + sx) * C) * esize;
        unsigned char* dst = out + (((long long)blockIdx.y * n_px + p) * C) * esize;
        for (int b = 0; b < C * esize; ++b) dst[b] = src[b];
    }
}
void launch_tile_raw(y3_context* ctx, const void* img_dev, int esize, long long row_lo, int W, int C,
                     const TileGeo* geo_dev, int count, int th, int tw, void* out) {
    if (count <= 0) return;
    dim3 grid((th * tw + 255) / 256 > 64 ? 64 : (th * tw + 255) / 256, count);
    k_tile_raw<<<grid, 256, 0, ctx->stream>>>(static_cast<const unsigned char*>(img_dev), esize, row_lo, W, C, geo_dev, th, tw,
                                              static_cast<unsigned char*>(out));
    Y3_LAUNCHED(ctx);
}

// ------------------------------------------------------------------------------------------ stitch
__global__ void __launch_bounds__(256)
k_stitch_flags(const float4* __restrict__ boxes, const int32_t* __restrict__ tile, int64_t n,
               const TileGeo* __restrict__ geo, StitchArgs S, int4* __restrict__ ibox, uint8_t* __restrict__ flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int4 ib;
    const bool keep = stitch_box(boxes[i], geo[tile[i]], S, &ib);
    ibox[i] = ib;
    flags[i] = keep ? 1 : 0;
}

__global__ void __launch_bounds__(CMP_BLOCK)
k_stitch_scatter(const uint8_t* __restrict__ flags, int64_t n, const int* __restrict__ blk, const int4* __restrict__ ibox,
                 const float* __restrict__ scores, const int32_t* __restrict__ labels, double* __restrict__ preds) {
    __shared__ int s_w[CMP_BLOCK / 32];
    const int64_t p = (int64_t)blockIdx.x * CMP_BLOCK + threadIdx.x;
    const bool f = p < n && flags[p];
    const int rank = block_rank(f, s_w);
    if (f) {
        double* o = preds + ((int64_t)blk[blockIdx.x] + rank) * 6;
        const int4 b = ibox[p];
        o[0] = b.x; o[1] = b.y; o[2] = b.z; o[3] = b.w;
        o[4] = (double)scores[p];
        o[5] = (double)labels[p];
    }
}

int64_t Tiler::stitch(PostProc* post, const NmsResult& R, const TileGeo* geo_dev, const StitchArgs& S) {
    if (R.n_kept <= 0) return 0;
    cudaStream_t st = ctx->stream;
    ibox.reserve((size_t)R.n_kept * 16);
    flags.reserve((size_t)R.n_kept);
    k_stitch_flags<<<ceil_div(R.n_kept, 256), 256, 0, st>>>(R.boxes, R.img, R.n_kept, geo_dev, S, ibox.as<int4>(),
                                                           flags.as<uint8_t>());
    Y3_LAUNCHED(ctx);
    const int64_t total = post->flag_offsets(flags.as<uint8_t>(), R.n_kept);
    if (total == 0) return 0;
    // grow the accumulator, preserving what earlier tile batches produced
    const size_t need = (size_t)(acc_rows + total) * 48;
    if (need > acc.cap) {
        DevBuf bigger;
        bigger.reserve(std::max(need, acc.cap * 2));
        if (acc_rows) Y3_CUDA(cudaMemcpyAsync(bigger.p, acc.p, (size_t)acc_rows * 48, cudaMemcpyDeviceToDevice, st));
        Y3_CUDA(cudaStreamSynchronize(st));
        acc.release();
        acc.p = bigger.p; acc.cap = bigger.cap;
        bigger.p = nullptr; bigger.cap = 0;
    }
    k_stitch_scatter<<<ceil_div(R.n_kept, CMP_BLOCK), CMP_BLOCK, 0, st>>>(flags.as<uint8_t>(), R.n_kept, post->blk.as<int>(),
                                                                         ibox.as<int4>(), R.scores, R.labels,
                                                                         acc.as<double>() + acc_rows * 6);
    Y3_LAUNCHED(ctx);
    acc_rows += total;
    return total;
}

// ------------------------------------------------------------------------------------------ cross-seam NMS
// Seam candidates: rows whose inclusive pixel extent straddles a zone boundary (zone = tile - 2*edge on every axis that
// is actually tiled, inference_tiled.py:41-47) - the boxes two neighbouring tiles can both have reported.
__global__ void __launch_bounds__(256)
k_seam_prepare(const double* __restrict__ preds, int64_t n, int zone_y, int zone_x, float4* __restrict__ box,
               float* __restrict__ score, int32_t* __restrict__ label, uint8_t* __restrict__ cand, uint8_t* __restrict__ keepm) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* p = preds + i * 6;
    const double x0 = p[0], y0 = p[1], x1 = p[2], y1 = p[3];
    bool c = false;
    if (zone_y > 0) c |= floor(y0 / zone_y) != floor(y1 / zone_y);
    if (zone_x > 0) c |= floor(x0 / zone_x) != floor(x1 / zone_x);
    box[i] = make_float4((float)x0, (float)y0, (float)x1, (float)y1);
    score[i] = (float)p[4];
    label[i] = (int32_t)p[5];
    cand[i] = c ? 1 : 0;
    keepm[i] = c ? 0 : 1;                       // non-candidates always stay
}
__global__ void __launch_bounds__(256)
k_seam_mark(const int32_t* __restrict__ src_row, int64_t n_kept, uint8_t* __restrict__ keepm) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_kept) keepm[src_row[i]] = 1;
}
__global__ void __launch_bounds__(CMP_BLOCK)
k_rows_scatter(const uint8_t* __restrict__ flags, int64_t n, const int* __restrict__ blk, const double* __restrict__ in,
               double* __restrict__ out) {
    __shared__ int s_w[CMP_BLOCK / 32];
    const int64_t p = (int64_t)blockIdx.x * CMP_BLOCK + threadIdx.x;
    const bool f = p < n && flags[p];
    const int rank = block_rank(f, s_w);
    if (f) {
        const double* src = in + p * 6;
        double* o = out + ((int64_t)blk[blockIdx.x] + rank) * 6;
#pragma unroll
        for (int k = 0; k < 6; ++k) o[k] = src[k];
    }
}

int64_t Tiler::cross_seam(PostProc* post, const double* preds_dev, int64_t n, const StitchArgs& S, int nc, float iou_thr) {
    if (n <= 0) return 0;
    cudaStream_t st = ctx->stream;
    seam_box.reserve((size_t)n * 16); seam_score.reserve((size_t)n * 4); seam_label.reserve((size_t)n * 4);
    seam_cand.reserve((size_t)n); seam_keep.reserve((size_t)n); seam_out.reserve((size_t)n * 48);
    const int zone_y = S.tile_h >= S.img_h ? 0 : S.tile_h - 2 * S.edge;
    const int zone_x = S.tile_w >= S.img_w ? 0 : S.tile_w - 2 * S.edge;
    k_seam_prepare<<<ceil_div(n, 256), 256, 0, st>>>(preds_dev, n, zone_y, zone_x, seam_box.as<float4>(), seam_score.as<float>(),
                                                     seam_label.as<int32_t>(), seam_cand.as<uint8_t>(), seam_keep.as<uint8_t>());
    Y3_LAUNCHED(ctx);
    static const bool use_grid = getenv("Y3_SEAM_SERIAL") == nullptr;
    const bool fast = use_grid && cross_seam_grid(seam_box.as<float4>(), seam_score.as<float>(), seam_label.as<int32_t>(),
                                                  seam_cand.as<uint8_t>(), n, S, iou_thr, seam_keep.as<uint8_t>());
    if (!fast) {
    CandSource src;
    src.box = seam_box.as<float>(); src.box_stride = 4;
    src.cls = seam_score.as<float>(); src.cls_stride = 1; src.obj = nullptr;
    src.rows_per_image = n; src.n_images = 1; src.nc = 1; src.raw_scores = true;
    src.row_seg = seam_label.as<int32_t>(); src.row_mask = seam_cand.as<uint8_t>(); src.n_seg_override = nc;
    const NmsResult R = post->run(src, iou_thr);
    if (R.n_kept > 0) {
        k_seam_mark<<<ceil_div(R.n_kept, 256), 256, 0, st>>>(R.src_row, R.n_kept, seam_keep.as<uint8_t>());
        Y3_LAUNCHED(ctx);
    }
    }
    const int64_t total = post->flag_offsets(seam_keep.as<uint8_t>(), n);
    if (total > 0) {
        k_rows_scatter<<<ceil_div(n, CMP_BLOCK), CMP_BLOCK, 0, st>>>(seam_keep.as<uint8_t>(), n, post->blk.as<int>(), preds_dev,
                                                                     seam_out.as<double>());
        Y3_LAUNCHED(ctx);
    }
    return total;
}

}  // namespace y3
