/* yolo3_b200.h - C ABI of libyolo3_b200.so: the B200-native (sm_100a) tiled YOLOv3 inference
 * hot path of usnistgov/object-detection-yolov3.
 *
 * The reference has NO FFI / plugin interface: its boundary is a set of plain Python call sites
 * (SURVEY.md section 8b).  Each entry point below names the reference symbol(s) it replaces
 * (file:line into the reference repo); the Python facades in object-detection-yolov3_b200/
 * (inference.py, inference_tiled.py, bbox_utils.py, model.py, imagereader.py) keep those
 * symbols' names and arguments and call these functions through ctypes.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types.
 *   - return 0 (Y3_OK) or a negative y3_status; y3_last_error() gives the message.  No
 *     exceptions or longjmp cross the boundary.
 *   - the caller allocates every output and passes its capacity; Y3_ERR_NOSPACE reports the
 *     required count through the same out-parameter that normally returns the count.
 *   - every call is synchronous on return (internal streams / graphs are hidden).
 *   - a handle is not thread-safe; distinct handles are independent.
 *   - pointers marked "host|device" are described by a y3_mem argument.  Device pointers must
 *     belong to the handle's CUDA device.
 *   - there is no CPU fallback: without a usable CUDA device every compute call fails loudly.
 */
#ifndef YOLO3_B200_H_
#define YOLO3_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define Y3_ABI_VERSION 3   /* 2: + y3_zscore; 3: + y3_comm_*, y3_infer_tiled_sharded, y3_cross_seam_nms, y3_timings.ms_comm */
#define Y3_MAX_ANCHORS 8

typedef int32_t y3_status;
enum {
    Y3_OK = 0,
    Y3_ERR_INVALID = -1,     /* bad argument */
    Y3_ERR_CUDA = -2,        /* CUDA runtime / driver failure (message has the detail) */
    Y3_ERR_NOSPACE = -3,     /* output capacity too small; required size returned */
    Y3_ERR_STATE = -4,       /* e.g. forward before weights were loaded */
    Y3_ERR_UNSUPPORTED = -5, /* shape the sm_100a kernels do not cover */
    Y3_ERR_NODEVICE = -6     /* no CUDA device / not an sm_100 part */
};

typedef enum { Y3_MEM_HOST = 0, Y3_MEM_DEVICE = 1 } y3_mem;
typedef enum { Y3_U8 = 0, Y3_U16 = 1, Y3_I32 = 2, Y3_F32 = 3 } y3_dtype;

typedef struct y3_context* y3_handle;

/* Mirrors model.YoloV3.__init__(global_batch_size, img_size[H,W,C], number_classes, anchors)
 * (model.py:423-451).  img_h == 0 creates a post-processing-only handle (NMS / tiling calls). */
typedef struct {
    int32_t struct_size;            /* sizeof(y3_config), for ABI evolution */
    int32_t img_h, img_w, img_c;    /* multiples of 32 (model.py:440, inference_tiled.py:38-39) */
    int32_t num_classes;
    int32_t num_anchors;            /* ALL anchors are used at every scale (model.py:108-111) */
    float anchors[Y3_MAX_ANCHORS][2]; /* (w, h) pixels; default (32,32),(128,128),(256,256) model.py:433 */
    int32_t max_batch;              /* largest number of images (tiles) per forward launch sequence; the tiled entry
                                     * points split a call's tiles evenly into batches of at most this many */
    int32_t device;                 /* CUDA ordinal (the reference uses CUDA_VISIBLE_DEVICES) */
    int64_t max_candidates;         /* capacity of the (box,class) candidate list; 0 = default */
} y3_config;

/* DLPack (v0.8 ABI) - weights are handed over as DLManagedTensor*; the library copies /
 * repacks and then calls each tensor's deleter. */
#ifndef DLPACK_DLPACK_H_
typedef enum { kDLCPU = 1, kDLCUDA = 2, kDLCUDAHost = 3 } DLDeviceType;
typedef struct { int32_t device_type; int32_t device_id; } DLDevice;
typedef struct { uint8_t code; uint8_t bits; uint16_t lanes; } DLDataType;
typedef struct {
    void* data; DLDevice device; int32_t ndim; DLDataType dtype;
    int64_t* shape; int64_t* strides; uint64_t byte_offset;
} DLTensor;
typedef struct DLManagedTensor {
    DLTensor dl_tensor; void* manager_ctx; void (*deleter)(struct DLManagedTensor* self);
} DLManagedTensor;
#endif

int32_t y3_abi_version(void);
/* message of the last failing call on this handle (or of the last failing y3_create when h == NULL) */
const char* y3_last_error(y3_handle h);

/* replaces: model.YoloV3(...) construction + tf.saved_model.load (model.py:423-464,
 * inference.py:35, inference_tiled.py:325). */
y3_status y3_create(const y3_config* cfg, y3_handle* out);
void y3_destroy(y3_handle h);

/* replaces: the Keras variables of the SavedModel.  names[i] are Keras variable names in creation
 * order ("conv2d_7/kernel", "conv2d_7/bias", "batch_normalization_7/gamma|beta|moving_mean|
 * moving_variance", "conv2d_transpose/kernel|bias", "feature_map_1/kernel|bias"); tensors are fp32
 * in Keras layouts (Conv2D kernel [kh,kw,Cin,Cout], Conv2DTranspose kernel [kh,kw,Cout,Cin]),
 * kDLCPU / kDLCUDAHost / kDLCUDA.  All 75 conv + 72 BN + 2 convT layers must be present. */
y3_status y3_load_weights(y3_handle h, int32_t n, const char* const* names,
                          DLManagedTensor* const* tensors);

/* replaces: YoloV3.get_keras_feature_map_model()(batch) (model.py:462,469) - the three raw heads,
 * NCHW fp32: fm1 [B,A(5+NC),H/32,W/32], fm2 [.. /16], fm3 [.. /8].  in: NCHW fp32 [B,C,H,W]. */
y3_status y3_forward_heads(y3_handle h, const float* in_nchw, y3_mem in_mem, int32_t batch,
                           float* fm1, float* fm2, float* fm3, y3_mem out_mem);

/* replaces: yolo_model(batch, training=False) (inference.py:58, inference_tiled.py:215;
 * model.py:169-212): decoded boxes [B, N, 5+NC] fp32, N = A*(g^2 + 4g^2 + 16g^2) rows ordered
 * scale 32,16,8 then (i*W_s + j)*A + a, columns x0,y0,x1,y1,obj,p_0.. (not clipped). */
y3_status y3_forward_boxes(y3_handle h, const float* in_nchw, y3_mem in_mem, int32_t batch,
                           float* out, y3_mem out_mem);
int64_t y3_boxes_per_image(y3_handle h);

/* replaces: yolo_model(...) -> filter_small_boxes -> per_class_nms for a batch of images
 * (inference.py:58-79; inference_tiled.py:215-228), everything on the device.
 * Outputs are image-major, then class-major, then score-descending (the reference's order).
 * img_index[k] says which image of the batch box k belongs to.  *n_out returns the count. */
y3_status y3_detect(y3_handle h, const float* in_nchw, y3_mem in_mem, int32_t batch,
                    float min_box_size, float iou_thr, float score_thr,
                    float* out_boxes /*[cap,4]*/, float* out_scores /*[cap]*/,
                    int32_t* out_labels /*[cap]*/, int32_t* out_img_index /*[cap]*/,
                    int64_t cap, int64_t* n_out);

/* replaces: the per-image body of inference.inference (inference.py:47-79) in ONE call: astype(float32) ->
 * imagereader.zscore_normalize over the whole image -> HWC->NCHW -> yolo_model(batch) -> clip of the corners to the image
 * (clip != 0; inference.py:62-65 as intended: x to [0, W], y to [0, H]) -> filter_small_boxes -> per_class_nms.
 * img: HWC, H x W x C = the network input size, any y3_dtype, host|device.  Outputs in the reference's order
 * (class-major, score-descending), not yet converted to x,y,w,h.  The batch-1 forward replays a CUDA graph. */
y3_status y3_detect_image(y3_handle h, const void* img, y3_dtype dtype, y3_mem img_mem, int32_t img_h, int32_t img_w, int32_t img_c,
                          float min_box_size, float iou_thr, float score_thr, int32_t clip,
                          float* out_boxes /*[cap,4]*/, float* out_scores /*[cap]*/, int32_t* out_labels /*[cap]*/,
                          int64_t cap, int64_t* n_out);

/* replaces: bbox_utils.compute_iou (bbox_utils.py:200-214 = inference_tiled.py:103-117).
 * iou[j] = IoU(box, boxes[j]) in the reference's fp32 operand order (0/0 -> NaN). */
y3_status y3_compute_iou(y3_handle h, const float* box /*[4]*/, const float* boxes /*[m,4]*/,
                         int64_t m, float* iou /*[m]*/);

/* replaces: bbox_utils.filter_small_boxes (bbox_utils.py:274-281 = inference_tiled.py:176-182).
 * rows [n, row_len] fp32 with x0,y0,x1,y1 first; keeps rows with (x1-x0) > min AND (y1-y0) > min,
 * order preserved.  out_rows [cap,row_len]; out_index (optional, may be NULL) the kept row ids. */
y3_status y3_filter_small(y3_handle h, const float* rows, int64_t n, int32_t row_len, float min_size,
                          float* out_rows, int64_t* out_index, int64_t cap, int64_t* n_out);

/* replaces: bbox_utils.single_class_nms (bbox_utils.py:217-237 = inference_tiled.py:120-140).
 * keep[] receives indices into boxes[] in pick (score-descending) order - the bit-exact target.
 * Ties: score descending, then index ascending (the reference's argsort()[::-1] is unpinned). */
y3_status y3_single_class_nms(y3_handle h, const float* boxes /*[m,4]*/, const float* scores /*[m]*/,
                              int64_t m, float iou_thr, int32_t* keep /*[m]*/, int64_t* n_keep);

/* replaces: bbox_utils.per_class_nms (bbox_utils.py:240-271 = inference_tiled.py:143-173).
 * score = sqrt(cls*obj) >= score_thr per class; outputs class-major, score-descending.
 * *n_out == 0 is the reference's (None, None, None). out_src (optional) = input row of each box. */
y3_status y3_per_class_nms(y3_handle h, const float* boxes /*[n,4]*/, const float* objectness /*[n]*/,
                           const float* class_probs /*[n,nc]*/, int64_t n, int32_t nc,
                           float iou_thr, float score_thr,
                           float* out_boxes /*[cap,4]*/, float* out_scores, int32_t* out_labels,
                           int32_t* out_src, int64_t cap, int64_t* n_out);

/* replaces: inference_tiled.convert_image_to_tiles geometry (inference_tiled.py:29-100).
 * Returns the tile count; xs/ys (optional, capacity cap) receive the RECORDED origins (the
 * clamped ones - the reference's border-tile quirk is reproduced on purpose). */
int64_t y3_tile_plan(int64_t img_h, int64_t img_w, int32_t tile_h, int32_t tile_w, int32_t edge_range,
                     int32_t* xs, int32_t* ys, int64_t cap);

/* How y3_infer_tiled / y3_infer_tiled_sharded batch `tile_count` tiles on a handle created with `max_batch`
 * (no reference counterpart: the reference runs one tile per model call, inference_tiled.py:213-216).  With the image
 * in host memory the first batch is short (<= 48 tiles: the convolutions start after one row of tiles has been
 * uploaded); the other tiles are split evenly into as few batches as max_batch allows, but into at least three
 * batches per call unless that would make them smaller than 32 tiles.  Returns the number of batches; sizes
 * (optional, capacity cap) receives their tile counts.  Host-only, needs no handle. */
int64_t y3_batch_plan(int64_t tile_count, int32_t max_batch, int32_t host_image, int32_t* sizes, int64_t cap);

/* replaces: inference_tiled.convert_image_to_tiles pixels (inference_tiled.py:29-100): the raw tiles
 * [first, first+count) in the SOURCE dtype, HWC each: out [count, tile_h, tile_w, C] (host|device). */
y3_status y3_tiles_raw(y3_handle h, const void* img, y3_dtype dtype, y3_mem img_mem,
                       int64_t img_h, int64_t img_w, int32_t img_c,
                       int32_t tile_h, int32_t tile_w, int32_t edge_range,
                       int64_t first, int64_t count, void* out, y3_mem out_mem);

/* replaces: convert_image_to_tiles + astype(float32) + imagereader.zscore_normalize + HWC->NCHW
 * (inference_tiled.py:29-100, 202-212; imagereader.py:34-46) for tiles [first, first+count):
 * out [count, C, tile_h, tile_w] fp32 (host|device).  img is HWC (host|device). */
y3_status y3_tiles_normalized(y3_handle h, const void* img, y3_dtype dtype, y3_mem img_mem,
                              int64_t img_h, int64_t img_w, int32_t img_c,
                              int32_t tile_h, int32_t tile_w, int32_t edge_range,
                              int64_t first, int64_t count, float* out, y3_mem out_mem);

/* replaces: imagereader.zscore_normalize (imagereader.py:34-46) on an array of ANY shape (no multiple-of-32
 * rule): out[i] = (x[i] - mean) / std with the population std over all n elements, x - mean when std <= 1.
 * data: n elements of `dtype` (host|device); out: n fp32 (host|device). */
y3_status y3_zscore(y3_handle h, const void* data, y3_dtype dtype, y3_mem data_mem, int64_t n, float* out, y3_mem out_mem);

/* replaces: the post-network part of inference_image_tiled (inference_tiled.py:218-310) with the
 * decoded boxes of every tile supplied by the caller - dets [count, N, 5+nc] fp32 (host|device):
 * filter_small_boxes -> per_class_nms -> ghost-band ownership -> origin add -> np.round ->
 * centre-in-image filter -> clamp.  preds [cap,6] float64 rows x0,y0,x1,y1,score,label in the
 * reference's order.  Tiles are [first, first+count) of the plan. */
y3_status y3_stitch_tiles(y3_handle h, const float* dets, y3_mem dets_mem, int64_t n_per_tile, int32_t nc,
                          int64_t img_h, int64_t img_w, int32_t tile_h, int32_t tile_w,
                          int32_t edge_range, int64_t first, int64_t count,
                          float min_box_size, float iou_thr, float score_thr,
                          double* preds, y3_mem preds_mem, int64_t cap, int64_t* n_out);

/* replaces: inference_tiled.inference_image_tiled (inference_tiled.py:185-310) for tiles
 * [tile_first, tile_first+tile_count) (tile_count < 0 = to the end) - the unit that is sharded
 * across GPUs.  img is the WHOLE HWC image (host|device); only the rows the tile range needs are
 * copied to the device (band by band on a copy stream, behind a short first batch, when the image is in
 * page-locked host memory).  The result does not depend on how the tiles are batched.  preds as in
 * y3_stitch_tiles; a host preds buffer is best page-locked (y3_host_alloc). */
y3_status y3_infer_tiled(y3_handle h, const void* img, y3_dtype dtype, y3_mem img_mem,
                         int64_t img_h, int64_t img_w, int32_t img_c,
                         int32_t tile_h, int32_t tile_w, int32_t edge_range,
                         int64_t tile_first, int64_t tile_count,
                         float min_box_size, float iou_thr, float score_thr,
                         double* preds, y3_mem preds_mem, int64_t cap, int64_t* n_out);

/* Page-locked host memory for images (pinned-host staging).  An image that lives in such a buffer
 * is uploaded by y3_infer_tiled band by band on a copy stream while the previous tile batch computes; a pageable image
 * works too, but its copies are staged synchronously by the driver.  imagereader.imread reads into these buffers. */
y3_status y3_host_alloc(int64_t bytes, void** out);
void y3_host_free(void* p);

/* Optional final stage of the tiled path that the reference does NOT have (it resolves seams by centre ownership only,
 * inference_tiled.py:235-254) and BASELINE.json's north_star asks for: among the boxes of preds [n,6] (float64 rows
 * x0,y0,x1,y1,score,label as returned by y3_infer_tiled, integer pixel corners) whose extent crosses a zone boundary of
 * the tile grid, greedy per-class NMS (same IoU arithmetic and tie rule as y3_single_class_nms) runs and suppressed rows
 * are dropped; all other rows and the row order are untouched.  nc = number of classes (labels are 0..nc-1). */
y3_status y3_cross_seam_nms(y3_handle h, const double* preds, y3_mem preds_mem, int64_t n, int32_t nc,
                            int64_t img_h, int64_t img_w, int32_t tile_h, int32_t tile_w, int32_t edge_range, float iou_thr,
                            double* out, y3_mem out_mem, int64_t cap, int64_t* n_out);

/* Multi-GPU (the reference is single-process): one process per GPU, one handle per process.
 * y3_comm_unique_id fills a 128-byte NCCL unique id on ONE rank; the host program broadcasts it (any transport) and
 * every rank calls y3_comm_init(h, rank, nranks, id) once.  NCCL is loaded at run time (libnccl.so.2). */
#define Y3_COMM_ID_BYTES 128
y3_status y3_comm_unique_id(uint8_t* id /*[Y3_COMM_ID_BYTES]*/);
y3_status y3_comm_init(y3_handle h, int32_t rank, int32_t nranks, const uint8_t* id);
int32_t y3_comm_size(y3_handle h);

/* replaces: inference_tiled.inference_image_tiled (inference_tiled.py:185-310) with the tile grid sharded across the
 * ranks of the communicator: every rank runs a contiguous row band of tiles (slice, normalise, network, decode, NMS,
 * ownership filter - no data-path collective), then the surviving rows are exchanged with ncclAllGather (counts, then
 * padded records) on the handle's stream and laid end to end in rank (= tile) order, i.e. every rank returns exactly
 * what y3_infer_tiled returns on one GPU.  img is the whole image in host memory or on THIS rank's device; only the
 * rows of the rank's band are uploaded.  cross_seam != 0 appends y3_cross_seam_nms over the gathered rows.  Collective:
 * all ranks must call it with the same arguments; a capacity overflow is reported by every rank together. */
y3_status y3_infer_tiled_sharded(y3_handle h, const void* img, y3_dtype dtype, y3_mem img_mem,
                                 int64_t img_h, int64_t img_w, int32_t img_c,
                                 int32_t tile_h, int32_t tile_w, int32_t edge_range,
                                 float min_box_size, float iou_thr, float score_thr, int32_t cross_seam,
                                 double* preds, y3_mem preds_mem, int64_t cap, int64_t* n_out);

/* Measurement hooks (not part of the reference surface): device time in ms of the stages of the
 * last y3_detect / y3_infer_tiled / NMS call, measured with CUDA events on the handle's stream,
 * and the number of kernels this library launched since the handle was created.  ms_decode is the fused
 * decode + threshold + small-box filter + compaction kernel alone (it is also contained in ms_nms). */
typedef struct {
    float ms_total, ms_h2d, ms_prep, ms_conv, ms_decode, ms_nms, ms_stitch, ms_d2h;
    float ms_comm;                  /* sharded path: count + record all-gathers, concatenation, optional cross-seam stage */
    float reserved_;
    int64_t kernels_launched;
    int64_t candidates, kept;
    int64_t tiles;                  /* tiles this handle processed in the last tiled call (its shard on the sharded path) */
} y3_timings;
y3_status y3_get_timings(y3_handle h, y3_timings* out);

/* Benchmark hook: run the conv stack only (no decode) `iters` times on the resident batch and
 * return the mean device ms per forward.  Used by bench.py for the roofline line. */
y3_status y3_bench_forward(y3_handle h, int32_t batch, int32_t iters, float* ms_per_iter);

/* Measurement hook: time every layer of the plan separately (CUDA events, `iters` repetitions each,
 * on whatever the activation buffers currently hold) and write a CSV report
 * "name,kind,k,stride,cin,cout,out_h,out_w,patch_h,patch_w,bn,bk,tiles,ms,tflops,min_gbytes_per_s"
 * into buf (NUL-terminated, truncated to cap). */
y3_status y3_profile_layers(y3_handle h, int32_t batch, int32_t iters, char* buf, int64_t cap);

/* Test hook (not part of the reference surface): the output of one layer of the LAST forward as
 * NCHW fp32 [batch, C, H, W]; layer is a Keras layer name ("conv2d_7", "conv2d_transpose").
 * dims receives C,H,W.  Meaningful for every layer only when the handle was created with the
 * environment variable Y3_DEBUG_NO_REUSE set (otherwise activation buffers are recycled). */
y3_status y3_debug_layer_output(y3_handle h, const char* layer, int32_t batch, float* out, int64_t cap_floats,
                                int32_t* dims /*[3]*/);

#ifdef __cplusplus
}
#endif
#endif /* YOLO3_B200_H_ */
