// tiles.cuh - tile geometry + front-end / back-end launchers (see tiles.cu).
#pragma once
#include "common.cuh"
#include "postproc.cuh"

namespace y3 {

std::vector<TileGeo> plan_tiles(int64_t H, int64_t W, int th, int tw, int edge, int* ry, int* rx);

void launch_tile_norm(y3_context* ctx, const void* img_dev, int dtype, long long row_lo, int W, int C,
                      const TileGeo* geo_dev, int count, int th, int tw, float* out, float* stats, double* sums_scratch);

void launch_tile_raw(y3_context* ctx, const void* img_dev, int esize, long long row_lo, int W, int C,
                     const TileGeo* geo_dev, int count, int th, int tw, void* out);

struct Tiler {
    y3_context* ctx;
    DevBuf geo, img, tiles, ibox, flags, acc, dets, sums;
    int64_t acc_rows = 0;
    float dbg_loop = 0.f;
    explicit Tiler(y3_context* c) : ctx(c) {}
    // appends the surviving boxes of R (image index = tile index inside geo_dev) to acc; returns how many
    int64_t stitch(PostProc* post, const NmsResult& R, const TileGeo* geo_dev, const StitchArgs& S);
};

}  // namespace y3
