// postproc.cuh - device-side post-processing pipeline (a13-a16 of SURVEY.md section 8):
// small-box filter + score threshold + compaction -> sort -> exact greedy per-class NMS.
#pragma once
#include "common.cuh"
#include "aux_kernels.cuh"

namespace y3 {

struct TileGeo {
    int y0, y1, x0, x1;     // clamped crop in the image
    int pre_y, pre_x;       // reflect padding before the crop
    int rec_x, rec_y;       // origin the reference records (clamped)
};

struct StitchArgs {
    int64_t img_h, img_w;
    int tile_h, tile_w, edge;
};

// Seam stitching folded into the NMS output stage of the tiled path (inference_tiled.py:235-301): a kept box is
// dropped unless its tile owns it, survivors are emitted as float64 [x0,y0,x1,y1,score,label] rows in image coordinates.
struct StitchCtx {
    const TileGeo* geo = nullptr;    // geometry of image (= tile) 0 of this batch
    StitchArgs S{};
    double* preds = nullptr;         // [cap_rows, 6] accumulated over the tile batches of one call
    int64_t cap_rows = 0;
};

// Device-resident control block of the synchronisation-free pipeline (nms_seg.cu).
struct PostCtrl {
    unsigned long long n_cand;       // candidates of this run (counted even beyond the capacity)
    int K;                           // candidates stored = min(n_cand, capacity); 0 when the run overflowed
    int overflow;                    // n_cand exceeded the candidate capacity
    int n_big, big_next;             // work list [0, n_big): segments one CTA resolves with the large key buffer; cursor
    int n_mid, mid_next;             // work list [n_big, n_mid): segments one CTA resolves (largest first); cursor
    int n_small, emit_ticket;        // work list [n_mid, n_small): segments one warp resolves (static round-robin); blocks of the fused emit done
    int max_seg;                     // largest segment
    int n_kept;                      // boxes this run emits (tiled: after the ownership filter)
    int n_kept_nms;                  // boxes NMS kept (before the ownership filter)
    int any_overflow;                // sticky over the batches of one call
    long long emit_base;             // tiled: first preds row of this run
    long long acc_rows;              // tiled: rows emitted so far by the batches of one call (counts beyond cap_rows too)
    long long sum_cand, sum_kept_nms;  // totals over the batches of one call (statistics)
};

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
    // Rows that carry their own segment id (the cross-seam stage: one list of boxes with a label each).  Needs
    // raw_scores, nc == 1, n_images == 1; row_mask (optional) selects the rows that take part at all.
    const int32_t* row_seg = nullptr;
    const uint8_t* row_mask = nullptr;
    int32_t n_seg_override = 0;
    int64_t num_segments() const { return row_seg ? (int64_t)n_seg_override : (int64_t)n_images * nc; }
#ifdef __CUDACC__
    __host__ __device__ int seg_image(int seg) const { return row_seg ? 0 : seg / nc; }
    __host__ __device__ int seg_label(int seg) const { return row_seg ? seg : seg % nc; }
#endif
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

static constexpr int SEG_MID_MAX = 24576;     // segmented pipeline: largest segment one CTA sorts in shared memory (192 KB of keys)

struct KeyLayout {
    int row_bits, seg_shift, total_bits;
    uint64_t row_mask;
    // score field = ~orderable(score) - score_base in score_bits bits: scores known to lie in [thr, 1] (the fused
    // sqrt(sigmoid*sigmoid) path) need 25 bits instead of 32 => one radix pass less
    uint32_t score_base, score_mask;
};

struct PostProc {
    y3_context* ctx;
    DevBuf keys[2], vals[2], sort_tmp, sbox, sarea, supp, keepf, seg_off, counters, blk, kbuf;
    DevBuf o_box, o_score, o_label, o_img, o_src, o_rank;
    // segmented pipeline (nms_seg.cu)
    DevBuf slot, seg_cnt, seg_off32, bkeys, rbox, rkey, kept_cnt, out_off, mid_list, ctrl, live, cbox, bbox;
    PinnedBuf host_small, host_ctrl;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float last_cand_ms = 0.f;    // device time of the last candidates launch of the global-sort path
    explicit PostProc(y3_context* c) : ctx(c) {}
    ~PostProc() {
        if (aux_stream) cudaStreamDestroy(aux_stream);
        for (cudaEvent_t e : {ev0, ev1, ev_fork, ev_join}) if (e) cudaEventDestroy(e);
    }
    // Plain result (boxes / scores / labels / image / source row on the device).  One host synchronisation (the
    // kept count) on the segmented path; the global-sort path for very large segments synchronises several times.
    NmsResult run(const CandSource& src, float iou_thr) { enqueue(src, iou_thr); return finish(); }
    // The same in two halves, so that a caller can stop its stage timer before the host waits: enqueue() launches
    // everything (the global-sort route completes inside it), finish() synchronises and returns the result.
    void enqueue(const CandSource& src, float iou_thr);
    NmsResult finish();
    // Tiled path: candidates -> per-segment NMS -> ownership filter -> float64 rows appended to st.preds, with NO host
    // synchronisation; counts live in the device control block until finish_tiled().  false = this source needs the
    // global-sort path (rows_per_image above the segmented pipeline's per-segment capacity): caller falls back to run().
    bool run_tiled(const CandSource& src, float iou_thr, const StitchCtx& st);
    void begin_tiled();                                   // resets the per-call accumulators (enqueued on ctx->stream)
    PostCtrl finish_tiled();                              // copies the control block to the host (synchronises ctx->stream)
    static bool segmented_ok(const CandSource& src);
    KeyLayout key_layout(const CandSource& src) const;
    int64_t capacity(const CandSource& src) const;
    void launch_candidates(const CandSource& src, const KeyLayout& kl, int64_t cap, bool count_segments);
    void segmented_front(const CandSource& src, const KeyLayout& kl, int64_t cap);     // memset + candidates + scan + bin
    void segmented_nms(const CandSource& src, const KeyLayout& kl, float iou_thr, const StitchCtx* st);
    void segmented_emit_plain(const CandSource& src, const KeyLayout& kl);
    NmsResult pending;           // result of a run that completed inside enqueue()
    int64_t pending_cap = -1;    // >= 0: a plain segmented run is in flight (its candidate capacity)
    bool fused_emit = false;     // the last segmented_nms left the output scan to k2_emit_fused
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    NmsResult run_global_sort(const CandSource& src, float iou_thr, const KeyLayout& kl, int64_t cap, int64_t K);
    // Ordered compaction support: exclusive per-1024-block offsets of set flags into `blk`
    // (int[ceil(n/1024)]), returns the total.  Synchronises the stream.
    int64_t flag_offsets(const uint8_t* flags, int64_t n);
};

#ifdef __CUDACC__
// inference_tiled.py:237-301 for one kept box of a tile (fp32 scalars, python-int operands converted to fp32):
// ghost-band ownership by box centre (:237-254), origin add (:263-266), np.round -> int32 (:278), centre inside the
// image (:281-288, int32 sum then / 2.0 in double), clamp to [0, size-1] (:291-301).  true = the box survives.
__device__ __forceinline__ bool stitch_box(const float4 b, const TileGeo g, const StitchArgs& S, int4* ibox) {
    const float r = (float)S.edge;
    const float ox = (float)g.rec_x, oy = (float)g.rec_y;
    const float cx = __fdiv_rn(__fadd_rn(b.z, b.x), 2.0f);
    const float cy = __fdiv_rn(__fadd_rn(b.w, b.y), 2.0f);
    const float gx = __fadd_rn(cx, ox);
    const float gy = __fadd_rn(cy, oy);
    bool bad = (gy > r) && (cy < r);
    bad |= (gy <= (float)(S.img_h - S.edge)) && (cy >= (float)(S.tile_h - S.edge));
    bad |= (gx > r) && (cx < r);
    bad |= (gx <= (float)(S.img_w - S.edge)) && (cx >= (float)(S.tile_w - S.edge));
    int x0 = (int)rintf(__fadd_rn(b.x, ox));
    int y0 = (int)rintf(__fadd_rn(b.y, oy));
    int x1 = (int)rintf(__fadd_rn(b.z, ox));
    int y1 = (int)rintf(__fadd_rn(b.w, oy));
    const double ccx = (double)(x1 + x0) / 2.0, ccy = (double)(y1 + y0) / 2.0;
    const bool outside = (ccx < 0) || (ccx >= (double)S.img_w) || (ccy < 0) || (ccy >= (double)S.img_h);
    const int mw = (int)S.img_w - 1, mh = (int)S.img_h - 1;
    x0 = min(max(x0, 0), mw); x1 = min(max(x1, 0), mw);
    y0 = min(max(y0, 0), mh); y1 = min(max(y1, 0), mh);
    *ibox = make_int4(x0, y0, x1, y1);
    return !bad && !outside;
}
#endif

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
