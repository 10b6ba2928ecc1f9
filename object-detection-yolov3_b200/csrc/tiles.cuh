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
    DevBuf seam_box, seam_score, seam_label, seam_cand, seam_keep, seam_out;      // cross-seam stage
    DevBuf shard_local, shard_gather, shard_counts;                               // sharded path (comm.cu)
    PinnedBuf grid_host;
    DevBuf grid_ctrl, grid_cells, grid_slot, grid_members, grid_state, grid_dom, grid_ndom;   // sparse parallel NMS (nms_grid.cu)
    DevBuf geo1;                                                                  // one "tile" = the whole image (y3_detect_image)
    bool geo1_ready = false;
    int geo1_h = 0, geo1_w = 0;
    int64_t acc_rows = 0;
    float dbg_loop = 0.f;
    explicit Tiler(y3_context* c) : ctx(c) {}
    // appends the surviving boxes of R (image index = tile index inside geo_dev) to acc; returns how many
    int64_t stitch(PostProc* post, const NmsResult& R, const TileGeo* geo_dev, const StitchArgs& S);
    // Optional final stage (north_star; not in the reference): greedy per-class NMS among the boxes whose extent crosses
    // a zone boundary of the tile grid; suppressed rows are dropped, order kept.  preds_dev [n,6] float64 on the device
    // -> seam_out (device), returns the surviving row count.  Synchronises the stream.
    int64_t cross_seam(PostProc* post, const double* preds_dev, int64_t n, const StitchArgs& S, int nc, float iou_thr);
    // sparse fast path of the stage (nms_grid.cu); false = not applicable, use the general pipeline
    bool cross_seam_grid(const float4* box, const float* score, const int32_t* label, const uint8_t* cand, int64_t n,
                         const StitchArgs& S, float iou_thr, uint8_t* keepm);
};

}  // namespace y3
