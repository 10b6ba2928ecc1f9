// postproc.cuh - device-side post-processing pipeline (a13-a16 of SURVEY.md section 8):
// small-box filter + score threshold + compaction -> sort -> exact greedy per-class NMS.
#pragma once
#include "common.cuh"
#include "aux_kernels.cuh"

namespace y3 {

// Describes decoded rows that already live on the device.
struct CandSource {
    const float* box = nullptr;  int64_t box_stride = 4;    // floats between rows; x0,y0,x1,y1
    const float* obj = nullptr;  int64_t obj_stride = 1;    // may be NULL => objectness 1 and score = cls
    const float* cls = nullptr;  int64_t cls_stride = 1;    // [rows, nc]
    int64_t rows_per_image = 0;
    int32_t n_images = 1;
    int32_t nc = 1;
    bool filter_small = false;   // fuse bbox_utils.filter_small_boxes
    float min_size = 0.f;
    bool raw_scores = false;     // score = cls (single_class_nms entry) instead of sqrt(cls*obj)
    float score_thr = 0.1f;      // ignored when raw_scores
    bool from_heads = false;     // fused path: decode on the fly from the raw fp32 heads (no decoded tensor)
    DecodeArgs dec;              // valid when from_heads
};

// Result lives in PostProc-owned device buffers until the next run().
struct NmsResult {
    int64_t n_cand = 0;
    int64_t n_kept = 0;
    const float4* boxes = nullptr;   // [n_kept]
    const float* scores = nullptr;
    const int32_t* labels = nullptr;
    const int32_t* img = nullptr;
    const int32_t* src_row = nullptr;  // row inside its image (index into the unfiltered rows)
};

struct PostProc {
    y3_context* ctx;
    DevBuf keys[2], vals[2], sort_tmp, sbox, sarea, supp, keepf, seg_off, counters, blk, kbuf;
    DevBuf o_box, o_score, o_label, o_img, o_src, o_rank;
    PinnedBuf host_small;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float last_cand_ms = 0.f;    // device time of the last k_candidates launch (decode+threshold+compaction)
    explicit PostProc(y3_context* c) : ctx(c) {}
    NmsResult run(const CandSource& src, float iou_thr);
    // Ordered compaction support: exclusive per-1024-block offsets of set flags into `blk`
    // (int[ceil(n/1024)]), returns the total.  Synchronises the stream.
    int64_t flag_offsets(const uint8_t* flags, int64_t n);
};

static constexpr int CMP_BLOCK = 1024;
// rank of this thread's set flag inside its 1024-thread block (s_w: int[32] shared scratch)
__device__ __forceinline__ int block_rank(bool f, int* s_w) {
    const unsigned b = __ballot_sync(0xffffffffu, f);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) s_w[wid] = __popc(b);
    __syncthreads();
    if (threadIdx.x < 32) {
        const int v = s_w[threadIdx.x];
        int incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (threadIdx.x >= o) incl += t;
        }
        s_w[threadIdx.x] = incl - v;
    }
    __syncthreads();
    return s_w[wid] + __popc(b & ((1u << lane) - 1u));
}

}  // namespace y3
