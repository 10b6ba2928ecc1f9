// api.cu - the extern "C" boundary declared in include/yolo3_b200.h.  Every entry point catches
// y3::Error (and anything else) and converts it into a negative status + message; nothing throws
// across the ABI.  There is no CPU fallback: without an sm_100 device y3_create fails.
#include "common.cuh"
#include "net.cuh"
#include "postproc.cuh"
#include "tiles.cuh"
#include "comm.cuh"

#include <algorithm>
#include <mutex>
#include <stdlib.h>
#include <time.h>

using namespace y3;

static thread_local std::string g_create_error;

#define Y3_API_BEGIN(h)                                                                   \
    if (!(h)) return Y3_ERR_INVALID;                                                      \
    try {                                                                                 \
        if (cudaSetDevice((h)->device) != cudaSuccess) fail(Y3_ERR_CUDA, "cudaSetDevice(%d) failed", (h)->device);
#define Y3_API_END(h)                                                                     \
        return Y3_OK;                                                                     \
    } catch (const Error& e) {                                                            \
        (h)->last_error = e.msg;                                                          \
        for (auto& r_ : (h)->phase_log) { (h)->event_pool.push_back(r_.a); (h)->event_pool.push_back(r_.b); } \
        (h)->phase_log.clear();                                                           \
        cudaGetLastError();                                                               \
        return e.code;                                                                    \
    } catch (const std::exception& e) {                                                   \
        (h)->last_error = std::string("internal error: ") + e.what();                     \
        return Y3_ERR_CUDA;                                                               \
    } catch (...) {                                                                       \
        (h)->last_error = "unknown internal error";                                       \
        return Y3_ERR_CUDA;                                                               \
    }

namespace {
// Sizes of the tile batches of one tiled call (see the comment in infer_tiled_impl; exported as y3_batch_plan).
std::vector<int> batch_plan(int64_t tile_count, int64_t max_batch, bool host_image) {
    std::vector<int> plan;
    int64_t left = tile_count;
    if (host_image && left > 0) { plan.push_back((int)std::min<int64_t>({max_batch, 48, left})); left -= plan.back(); }
    if (left > 0) {
        int64_t nb = (left + max_batch - 1) / max_batch;
        const int64_t want = 3 - (int64_t)plan.size();
        if (nb < want) nb = std::max(nb, std::min<int64_t>(want, left / 32));
        for (int64_t i = 0; i < nb; ++i) plan.push_back((int)(left / nb + (i < left % nb ? 1 : 0)));
    }
    return plan;
}


const void* to_device(y3_context* c, const void* p, y3_mem mem, size_t bytes, DevBuf& stage) {
    if (mem == Y3_MEM_DEVICE) return p;
    stage.reserve(bytes);
    Y3_CUDA(cudaMemcpyAsync(stage.p, p, bytes, cudaMemcpyHostToDevice, c->stream));
    return stage.p;
}
void from_device(y3_context* c, void* dst, y3_mem mem, const void* src, size_t bytes) {
    if (!bytes) return;
    Y3_CUDA(cudaMemcpyAsync(dst, src, bytes, mem == Y3_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
}
PostProc* post_of(y3_context* c) {
    if (!c->post) c->post = new PostProc(c);
    return c->post;
}
Tiler* tiler_of(y3_context* c) {
    if (!c->tiler) c->tiler = new Tiler(c);
    return c->tiler;
}
Net* net_of(y3_context* c) {
    Y3_CHECK(c->net, Y3_ERR_STATE, "this handle was created without a network (img_h == 0)");
    return c->net;
}
size_t dtype_size(int dt) { return dt == Y3_U8 ? 1 : dt == Y3_U16 ? 2 : 4; }

__global__ void k_iou_row(float4 box, const float* __restrict__ boxes, int64_t m, float* __restrict__ out) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const float4 b = make_float4(boxes[4 * j], boxes[4 * j + 1], boxes[4 * j + 2], boxes[4 * j + 3]);
    out[j] = iou_exact(box, box_area_exact(box), b, box_area_exact(b));
}
__global__ void k_small_flags(const float* __restrict__ rows, int64_t n, int row_len, float min_size, uint8_t* __restrict__ flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* r = rows + i * row_len;
    const float w = __fsub_rn(r[2], r[0]), h = __fsub_rn(r[3], r[1]);
    flags[i] = (w > min_size && h > min_size) ? 1 : 0;
}
__global__ void __launch_bounds__(CMP_BLOCK)
k_small_scatter(const float* __restrict__ rows, int64_t n, int row_len, const uint8_t* __restrict__ flags,
                const int* __restrict__ blk, float* __restrict__ out, long long* __restrict__ out_idx) {
    __shared__ int s_w[CMP_BLOCK / 32];
    __shared__ long long s_dst[CMP_BLOCK];
    const int64_t p = (int64_t)blockIdx.x * CMP_BLOCK + threadIdx.x;
    const bool f = p < n && flags[p];
    const int rank = block_rank(f, s_w);
    const long long q = (long long)blk[blockIdx.x] + rank;
    s_dst[threadIdx.x] = f ? q : -1;
    if (f && out_idx) out_idx[q] = p;
    __syncthreads();
    // copy rows cooperatively: consecutive threads copy consecutive floats
    const int64_t base = (int64_t)blockIdx.x * CMP_BLOCK;
    const int64_t cnt = min((int64_t)CMP_BLOCK, n - base);
    for (int64_t e = threadIdx.x; e < cnt * row_len; e += CMP_BLOCK) {
        const int r = (int)(e / row_len), k = (int)(e - (int64_t)r * row_len);
        const long long d = s_dst[r];
        if (d >= 0) out[d * row_len + k] = rows[(base + r) * row_len + k];
    }
}
// shared by y3_detect / y3_stitch_tiles / y3_infer_tiled
CandSource dets_source(const float* dets_dev, int64_t n_per_img, int n_img, int nc, bool filter, float min_box, float score_thr) {
    CandSource s;
    const int E = 5 + nc;
    s.box = dets_dev; s.box_stride = E;
    s.obj = dets_dev + 4; s.obj_stride = E;
    s.cls = dets_dev + 5; s.cls_stride = E;
    s.rows_per_image = n_per_img; s.n_images = n_img; s.nc = nc;
    s.filter_small = filter; s.min_size = min_box; s.score_thr = score_thr;
    return s;
}

// fused: the candidates are thresholded straight from the raw heads of the last forward
CandSource heads_source(const Net* net, int n_img, bool filter, float min_box, float score_thr, int head_set = 0, bool clip = false) {
    CandSource s;
    s.from_heads = true;
    s.dec = net->decode_args(n_img, head_set);
    if (clip) { s.dec.clip_w = (float)net->W; s.dec.clip_h = (float)net->H; }
    s.rows_per_image = net->rows_per_image; s.n_images = n_img; s.nc = net->nc;
    s.filter_small = filter; s.min_size = min_box; s.score_thr = score_thr;
    return s;
}

}  // namespace

// ============================================================================================
extern "C" {

int32_t y3_abi_version(void) { return Y3_ABI_VERSION; }

const char* y3_last_error(y3_handle h) { return h ? h->last_error.c_str() : g_create_error.c_str(); }

y3_status y3_create(const y3_config* cfg, y3_handle* out) {
    if (!cfg || !out) { g_create_error = "y3_create: NULL argument"; return Y3_ERR_INVALID; }
    *out = nullptr;
    y3_context* c = nullptr;
    try {
        Y3_CHECK(cfg->struct_size == (int32_t)sizeof(y3_config), Y3_ERR_INVALID, "y3_config.struct_size %d != %d (ABI mismatch)",
                 cfg->struct_size, (int)sizeof(y3_config));
        int n_dev = 0;
        cudaError_t e = cudaGetDeviceCount(&n_dev);
        if (e != cudaSuccess || n_dev == 0) {
            cudaGetLastError();
            fail(Y3_ERR_NODEVICE, "no CUDA device available (%s) - libyolo3_b200 has no CPU fallback",
                 e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        }
        Y3_CHECK(cfg->device >= 0 && cfg->device < n_dev, Y3_ERR_INVALID, "device %d outside 0..%d", cfg->device, n_dev - 1);
        cudaDeviceProp prop;
        Y3_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
        Y3_CHECK(prop.major == 10, Y3_ERR_NODEVICE, "device %d (%s, sm_%d%d) is not an sm_100 part - this library is B200-only",
                 cfg->device, prop.name, prop.major, prop.minor);
        Y3_CUDA(cudaSetDevice(cfg->device));
        c = new y3_context();
        c->cfg = *cfg;
        c->device = cfg->device;
        c->sm_count = prop.multiProcessorCount;
        Y3_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        if (cfg->img_h > 0) {
            c->net = new Net(c);
            c->net->build();
            Y3_CUDA(cudaDeviceSynchronize());
        }
        *out = c;
        return Y3_OK;
    } catch (const Error& e) {
        g_create_error = e.msg;
        if (c) y3_destroy(c);
        cudaGetLastError();
        return e.code;
    } catch (...) {
        g_create_error = "unknown internal error in y3_create";
        if (c) y3_destroy(c);
        return Y3_ERR_CUDA;
    }
}

void y3_destroy(y3_handle h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->comm) { try { comm_destroy(h); } catch (...) {} }
    delete h->net;
    delete h->post;
    delete h->tiler;
    for (cudaEvent_t e : h->event_pool) cudaEventDestroy(e);
    for (auto& r : h->phase_log) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->post_stream) cudaStreamDestroy(h->post_stream);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

y3_status y3_load_weights(y3_handle h, int32_t n, const char* const* names, DLManagedTensor* const* tensors) {
    Y3_API_BEGIN(h)
    Y3_CHECK(n >= 0 && (n == 0 || (names && tensors)), Y3_ERR_INVALID, "bad weight list");
    net_of(h)->load(n, names, tensors);
    Y3_API_END(h)
}

int64_t y3_boxes_per_image(y3_handle h) { return (h && h->net) ? h->net->rows_per_image : 0; }

y3_status y3_forward_heads(y3_handle h, const float* in, y3_mem in_mem, int32_t batch, float* fm1, float* fm2, float* fm3,
                           y3_mem out_mem) {
    Y3_API_BEGIN(h)
    Net* net = net_of(h);
    Y3_CHECK(in && fm1 && fm2 && fm3, Y3_ERR_INVALID, "NULL pointer");
    Y3_CHECK(batch >= 1 && batch <= net->maxB, Y3_ERR_INVALID, "batch %d outside 1..%d", batch, net->maxB);
    const size_t in_bytes = (size_t)batch * net->C * net->H * net->W * 4;
    const float* d_in = static_cast<const float*>(to_device(h, in, in_mem, in_bytes, h->stage_in));
    net->forward(d_in, batch);
    float* outs[3] = {fm1, fm2, fm3};
    for (int s = 0; s < 3; ++s) {
        const int hw = net->gh[s] * net->gw[s];
        const size_t bytes = (size_t)batch * net->det_c * hw * 4;
        float* dst = out_mem == Y3_MEM_DEVICE ? outs[s] : (h->stage_out.reserve(bytes), h->stage_out.as<float>());
        heads_to_nchw(h, net->head[s], dst, batch, hw, net->det_c, net->head_pitch, net->na, net->nc);
        if (out_mem != Y3_MEM_DEVICE) {
            from_device(h, outs[s], out_mem, dst, bytes);
            Y3_CUDA(cudaStreamSynchronize(h->stream));
        }
    }
    Y3_CUDA(cudaStreamSynchronize(h->stream));
    Y3_API_END(h)
}

y3_status y3_forward_boxes(y3_handle h, const float* in, y3_mem in_mem, int32_t batch, float* out, y3_mem out_mem) {
    Y3_API_BEGIN(h)
    Net* net = net_of(h);
    Y3_CHECK(in && out, Y3_ERR_INVALID, "NULL pointer");
    Y3_CHECK(batch >= 1 && batch <= net->maxB, Y3_ERR_INVALID, "batch %d outside 1..%d", batch, net->maxB);
    const size_t in_bytes = (size_t)batch * net->C * net->H * net->W * 4;
    const float* d_in = static_cast<const float*>(to_device(h, in, in_mem, in_bytes, h->stage_in));
    net->forward(d_in, batch);
    net->decode(batch);
    from_device(h, out, out_mem, net->boxes.p, (size_t)batch * net->rows_per_image * (5 + net->nc) * 4);
    Y3_CUDA(cudaStreamSynchronize(h->stream));
    Y3_API_END(h)
}

y3_status y3_detect(y3_handle h, const float* in, y3_mem in_mem, int32_t batch, float min_box, float iou_thr, float score_thr,
                    float* out_boxes, float* out_scores, int32_t* out_labels, int32_t* out_img, int64_t cap, int64_t* n_out) {
    Y3_API_BEGIN(h)
    Net* net = net_of(h);
    Y3_CHECK(in && n_out, Y3_ERR_INVALID, "NULL pointer");
    Y3_CHECK(batch >= 1 && batch <= net->maxB, Y3_ERR_INVALID, "batch %d outside 1..%d", batch, net->maxB);
    y3_timings& T = h->timings;
    T = y3_timings{};
    Phase total(h, &T.ms_total, "y3:total");
    const size_t in_bytes = (size_t)batch * net->C * net->H * net->W * 4;
    const float* d_in;
    { Phase p(h, &T.ms_h2d, "y3:h2d"); d_in = static_cast<const float*>(to_device(h, in, in_mem, in_bytes, h->stage_in)); p.stop(); }
    { Phase p(h, &T.ms_conv, "y3:conv_stack"); net->forward(d_in, batch); p.stop(); }
    NmsResult R;
    {   // decode + score + threshold + small-box filter + compaction fused in one pass over the heads
        Phase p(h, &T.ms_nms, "y3:decode_nms_stitch");
        post_of(h)->enqueue(heads_source(net, batch, true, min_box, score_thr), iou_thr);
        p.stop();                                    // (ms_decode: the fused decode+threshold+compaction kernel alone)
        R = post_of(h)->finish();
    }
    T.candidates = R.n_cand; T.kept = R.n_kept;
    *n_out = R.n_kept;
    Y3_CHECK(R.n_kept <= cap, Y3_ERR_NOSPACE, "output capacity %lld < %lld kept boxes", (long long)cap, (long long)R.n_kept);
    if (R.n_kept) {
        Phase p(h, &T.ms_d2h, "y3:d2h");
        Y3_CHECK(out_boxes && out_scores && out_labels && out_img, Y3_ERR_INVALID, "NULL output");
        Y3_CUDA(cudaMemcpyAsync(out_boxes, R.boxes, (size_t)R.n_kept * 16, cudaMemcpyDeviceToHost, h->stream));
        Y3_CUDA(cudaMemcpyAsync(out_scores, R.scores, (size_t)R.n_kept * 4, cudaMemcpyDeviceToHost, h->stream));
        Y3_CUDA(cudaMemcpyAsync(out_labels, R.labels, (size_t)R.n_kept * 4, cudaMemcpyDeviceToHost, h->stream));
        Y3_CUDA(cudaMemcpyAsync(out_img, R.img, (size_t)R.n_kept * 4, cudaMemcpyDeviceToHost, h->stream));
        p.stop();
    }
    total.stop();
    flush_phases(h);
    T.kernels_launched = h->kernels_launched;
    Y3_API_END(h)
}

y3_status y3_detect_image(y3_handle h, const void* img, y3_dtype dt, y3_mem img_mem, int32_t H, int32_t W, int32_t C, float min_box,
                          float iou_thr, float score_thr, int32_t clip, float* out_boxes, float* out_scores, int32_t* out_labels,
                          int64_t cap, int64_t* n_out) {
    Y3_API_BEGIN(h)
    Net* net = net_of(h);
    Tiler* T = tiler_of(h);
    Y3_CHECK(img && n_out, Y3_ERR_INVALID, "NULL pointer");
    Y3_CHECK(H == net->H && W == net->W && C == net->C, Y3_ERR_INVALID, "image %dx%dx%d does not match the network input %dx%dx%d",
             H, W, C, net->H, net->W, net->C);
    y3_timings& Tm = h->timings;
    Tm = y3_timings{};
    Phase total(h, &Tm.ms_total, "y3:total");
    const size_t in_bytes = (size_t)H * W * C * dtype_size(dt);
    const void* d_img;
    { Phase p(h, &Tm.ms_h2d, "y3:h2d"); d_img = to_device(h, img, img_mem, in_bytes, T->img); p.stop(); }
    {   // whole-image z-score (imagereader.py:34-46) + HWC -> NCHW: the tile front-end with one tile = the image
        Phase p(h, &Tm.ms_prep, "y3:tile_slice_zscore");
        if (!T->geo1_ready || T->geo1_h != H || T->geo1_w != W) {
            TileGeo g{};
            g.y0 = 0; g.y1 = H; g.x0 = 0; g.x1 = W;
            T->geo1.reserve(sizeof(TileGeo));
            Y3_CUDA(cudaMemcpyAsync(T->geo1.p, &g, sizeof(TileGeo), cudaMemcpyHostToDevice, h->stream));
            Y3_CUDA(cudaStreamSynchronize(h->stream));         // g is a host temporary
            T->geo1_ready = true; T->geo1_h = H; T->geo1_w = W;
        }
        T->tiles.reserve((size_t)net->maxB * C * H * W * 4);
        T->sums.reserve(16 * (size_t)net->maxB);
        launch_tile_norm(h, d_img, dt, 0, W, C, T->geo1.as<TileGeo>(), 1, H, W, T->tiles.as<float>(), nullptr, T->sums.as<double>());
        p.stop();
    }
    { Phase p(h, &Tm.ms_conv, "y3:conv_stack"); net->forward(T->tiles.as<float>(), 1); p.stop(); }
    NmsResult R;
    {
        Phase p(h, &Tm.ms_nms, "y3:decode_nms_stitch");
        post_of(h)->enqueue(heads_source(net, 1, true, min_box, score_thr, 0, clip != 0), iou_thr);
        p.stop();
        R = post_of(h)->finish();
    }
    Tm.candidates = R.n_cand; Tm.kept = R.n_kept;
    *n_out = R.n_kept;
    Y3_CHECK(R.n_kept <= cap, Y3_ERR_NOSPACE, "output capacity %lld < %lld kept boxes", (long long)cap, (long long)R.n_kept);
    if (R.n_kept) {
        Phase p(h, &Tm.ms_d2h, "y3:d2h");
        Y3_CHECK(out_boxes && out_scores && out_labels, Y3_ERR_INVALID, "NULL output");
        Y3_CUDA(cudaMemcpyAsync(out_boxes, R.boxes, (size_t)R.n_kept * 16, cudaMemcpyDeviceToHost, h->stream));
        Y3_CUDA(cudaMemcpyAsync(out_scores, R.scores, (size_t)R.n_kept * 4, cudaMemcpyDeviceToHost, h->stream));
        Y3_CUDA(cudaMemcpyAsync(out_labels, R.labels, (size_t)R.n_kept * 4, cudaMemcpyDeviceToHost, h->stream));
        p.stop();
    }
    total.stop();
    flush_phases(h);
    Tm.kernels_launched = h->kernels_launched;
    Y3_API_END(h)
}

y3_status y3_compute_iou(y3_handle h, const float* box, const float* boxes, int64_t m, float* iou) {
    Y3_API_BEGIN(h)
    Y3_CHECK(box && (m == 0 || (boxes && iou)) && m >= 0, Y3_ERR_INVALID, "bad arguments");
    if (m > 0) {
        const float* d = static_cast<const float*>(to_device(h, boxes, Y3_MEM_HOST, (size_t)m * 16, h->stage_in));
        h->stage_out.reserve((size_t)m * 4);
        k_iou_row<<<ceil_div(m, 256), 256, 0, h->stream>>>(make_float4(box[0], box[1], box[2], box[3]), d, m, h->stage_out.as<float>());
        Y3_LAUNCHED(h);
        from_device(h, iou, Y3_MEM_HOST, h->stage_out.p, (size_t)m * 4);
        Y3_CUDA(cudaStreamSynchronize(h->stream));
    }
    Y3_API_END(h)
}

y3_status y3_filter_small(y3_handle h, const float* rows, int64_t n, int32_t row_len, float min_size, float* out_rows,
                          int64_t* out_index, int64_t cap, int64_t* n_out) {
    Y3_API_BEGIN(h)
    Y3_CHECK(n_out && n >= 0 && row_len >= 4 && (n == 0 || rows), Y3_ERR_INVALID, "bad arguments");
    *n_out = 0;
    if (n > 0) {
        PostProc* P = post_of(h);
        const float* d = static_cast<const float*>(to_device(h, rows, Y3_MEM_HOST, (size_t)n * row_len * 4, h->stage_in));
        h->stage_aux.reserve((size_t)n);
        k_small_flags<<<ceil_div(n, 256), 256, 0, h->stream>>>(d, n, row_len, min_size, h->stage_aux.as<uint8_t>());
        Y3_LAUNCHED(h);
        const int64_t k = P->flag_offsets(h->stage_aux.as<uint8_t>(), n);
        *n_out = k;
        Y3_CHECK(k <= cap, Y3_ERR_NOSPACE, "output capacity %lld < %lld rows", (long long)cap, (long long)k);
        if (k > 0) {
            Y3_CHECK(out_rows, Y3_ERR_INVALID, "NULL output");
            const size_t idx_off = ((size_t)k * row_len * 4 + 7) & ~size_t(7);
            h->stage_out.reserve(idx_off + (size_t)k * 8);
            long long* d_idx = reinterpret_cast<long long*>(h->stage_out.as<unsigned char>() + idx_off);
            k_small_scatter<<<ceil_div(n, CMP_BLOCK), CMP_BLOCK, 0, h->stream>>>(d, n, row_len, h->stage_aux.as<uint8_t>(),
                                                                               P->blk.as<int>(), h->stage_out.as<float>(), d_idx);
            Y3_LAUNCHED(h);
            from_device(h, out_rows, Y3_MEM_HOST, h->stage_out.p, (size_t)k * row_len * 4);
            if (out_index) from_device(h, out_index, Y3_MEM_HOST, d_idx, (size_t)k * 8);
            Y3_CUDA(cudaStreamSynchronize(h->stream));
        }
    }
    Y3_API_END(h)
}

y3_status y3_single_class_nms(y3_handle h, const float* boxes, const float* scores, int64_t m, float iou_thr, int32_t* keep,
                              int64_t* n_keep) {
    Y3_API_BEGIN(h)
    Y3_CHECK(n_keep && m >= 0 && (m == 0 || (boxes && scores && keep)), Y3_ERR_INVALID, "bad arguments");
    *n_keep = 0;
    if (m > 0) {
        y3_timings& T = h->timings;
        T = y3_timings{};
        Phase total(h, &T.ms_total, "y3:total");
        h->stage_in.reserve((size_t)m * 20);
        float* d_box = h->stage_in.as<float>();
        float* d_sc = d_box + 4 * m;
        { Phase p(h, &T.ms_h2d, "y3:h2d");
          Y3_CUDA(cudaMemcpyAsync(d_box, boxes, (size_t)m * 16, cudaMemcpyHostToDevice, h->stream));
          Y3_CUDA(cudaMemcpyAsync(d_sc, scores, (size_t)m * 4, cudaMemcpyHostToDevice, h->stream));
          p.stop(); }
        CandSource s;
        s.box = d_box; s.box_stride = 4; s.cls = d_sc; s.cls_stride = 1; s.obj = nullptr;
        s.rows_per_image = m; s.n_images = 1; s.nc = 1; s.raw_scores = true;
        NmsResult R;
        { Phase p(h, &T.ms_nms, "y3:decode_nms_stitch"); post_of(h)->enqueue(s, iou_thr); p.stop(); R = post_of(h)->finish(); }
        *n_keep = R.n_kept;
        T.candidates = R.n_cand; T.kept = R.n_kept;
        { Phase p(h, &T.ms_d2h, "y3:d2h");
          from_device(h, keep, Y3_MEM_HOST, R.src_row, (size_t)R.n_kept * 4);
          p.stop(); }
        total.stop();
        flush_phases(h);
        T.kernels_launched = h->kernels_launched;
    }
    Y3_API_END(h)
}

y3_status y3_per_class_nms(y3_handle h, const float* boxes, const float* obj, const float* cls, int64_t n, int32_t nc,
                           float iou_thr, float score_thr, float* out_boxes, float* out_scores, int32_t* out_labels,
                           int32_t* out_src, int64_t cap, int64_t* n_out) {
    Y3_API_BEGIN(h)
    Y3_CHECK(n_out && n >= 0 && nc >= 1 && (n == 0 || (boxes && obj && cls)), Y3_ERR_INVALID, "bad arguments");
    *n_out = 0;
    if (n > 0) {
        y3_timings& T = h->timings;
        T = y3_timings{};
        Phase total(h, &T.ms_total, "y3:total");
        h->stage_in.reserve((size_t)n * (5 + nc) * 4);
        float* d_box = h->stage_in.as<float>();
        float* d_obj = d_box + 4 * n;
        float* d_cls = d_obj + n;
        { Phase p(h, &T.ms_h2d, "y3:h2d");
          Y3_CUDA(cudaMemcpyAsync(d_box, boxes, (size_t)n * 16, cudaMemcpyHostToDevice, h->stream));
          Y3_CUDA(cudaMemcpyAsync(d_obj, obj, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
          Y3_CUDA(cudaMemcpyAsync(d_cls, cls, (size_t)n * nc * 4, cudaMemcpyHostToDevice, h->stream));
          p.stop(); }
        CandSource s;
        s.box = d_box; s.box_stride = 4; s.obj = d_obj; s.obj_stride = 1; s.cls = d_cls; s.cls_stride = nc;
        s.rows_per_image = n; s.n_images = 1; s.nc = nc; s.score_thr = score_thr;
        NmsResult R;
        { Phase p(h, &T.ms_nms, "y3:decode_nms_stitch"); post_of(h)->enqueue(s, iou_thr); p.stop(); R = post_of(h)->finish(); }
        *n_out = R.n_kept;
        T.candidates = R.n_cand; T.kept = R.n_kept;
        Y3_CHECK(R.n_kept <= cap, Y3_ERR_NOSPACE, "output capacity %lld < %lld kept boxes", (long long)cap, (long long)R.n_kept);
        if (R.n_kept) {
            Phase p(h, &T.ms_d2h, "y3:d2h");
            Y3_CHECK(out_boxes && out_scores && out_labels, Y3_ERR_INVALID, "NULL output");
            from_device(h, out_boxes, Y3_MEM_HOST, R.boxes, (size_t)R.n_kept * 16);
            from_device(h, out_scores, Y3_MEM_HOST, R.scores, (size_t)R.n_kept * 4);
            from_device(h, out_labels, Y3_MEM_HOST, R.labels, (size_t)R.n_kept * 4);
            if (out_src) from_device(h, out_src, Y3_MEM_HOST, R.src_row, (size_t)R.n_kept * 4);
            p.stop();
        }
        total.stop();
        flush_phases(h);
        T.kernels_launched = h->kernels_launched;
    }
    Y3_API_END(h)
}

int64_t y3_tile_plan(int64_t img_h, int64_t img_w, int32_t tile_h, int32_t tile_w, int32_t edge, int32_t* xs, int32_t* ys, int64_t cap) {
    try {
        std::vector<TileGeo> v = plan_tiles(img_h, img_w, tile_h, tile_w, edge, nullptr, nullptr);
        for (size_t i = 0; i < v.size() && (int64_t)i < cap; ++i) {
            if (xs) xs[i] = v[i].rec_x;
            if (ys) ys[i] = v[i].rec_y;
        }
        return (int64_t)v.size();
    } catch (const Error& e) {
        g_create_error = e.msg;
        return e.code;
    }
}

int64_t y3_batch_plan(int64_t tile_count, int32_t max_batch, int32_t host_image, int32_t* sizes, int64_t cap) {
    if (tile_count < 0 || max_batch < 1) { g_create_error = "y3_batch_plan: bad arguments"; return Y3_ERR_INVALID; }
    const std::vector<int> v = batch_plan(tile_count, max_batch, host_image != 0);
    for (size_t i = 0; i < v.size() && (int64_t)i < cap; ++i)
        if (sizes) sizes[i] = v[i];
    return (int64_t)v.size();
}

namespace {
// uploads the rows of the image that tiles [first, first+count) touch; returns the device pointer and row_lo
const void* upload_band(y3_context* h, Tiler* T, const void* img, y3_dtype dt, y3_mem mem, int64_t W, int C,
                        const std::vector<TileGeo>& geo, int64_t first, int64_t count, long long* row_lo) {
    if (mem == Y3_MEM_DEVICE) { *row_lo = 0; return img; }
    int lo = geo[first].y0, hi = geo[first].y1;
    for (int64_t t = first; t < first + count; ++t) { lo = std::min(lo, geo[t].y0); hi = std::max(hi, geo[t].y1); }
    const size_t row_bytes = (size_t)W * C * dtype_size(dt);
    T->img.reserve((size_t)(hi - lo) * row_bytes);
    Y3_CUDA(cudaMemcpyAsync(T->img.p, static_cast<const char*>(img) + (size_t)lo * row_bytes, (size_t)(hi - lo) * row_bytes,
                            cudaMemcpyHostToDevice, h->stream));
    *row_lo = lo;
    return T->img.p;
}

// Same band, but uploaded just in time: the rows batch k+1 needs are copied on a separate stream while
// batch k computes (pinned host memory overlaps fully).  Small per-batch copies never hog a copy engine,
// so stream-ordered work of the compute stream is not queued behind a long transfer.
struct BandUpload {
    const void* dev = nullptr;      // device address of image row `row_lo`
    const char* host = nullptr;
    long long row_lo = 0;
    int hi = 0, next_row = 0;       // rows [row_lo, next_row) have been enqueued
    size_t row_bytes = 0;
    std::vector<cudaEvent_t> ev;    // events to return to the pool
};
BandUpload begin_band_upload(y3_context* h, Tiler* T, const void* img, y3_dtype dt, y3_mem mem, int64_t W, int C,
                             const std::vector<TileGeo>& geo, int64_t first, int64_t count) {
    BandUpload U;
    if (mem == Y3_MEM_DEVICE) { U.dev = img; return U; }
    int lo = geo[first].y0, hi = geo[first].y1;
    for (int64_t t = first; t < first + count; ++t) { lo = std::min(lo, geo[t].y0); hi = std::max(hi, geo[t].y1); }
    U.row_bytes = (size_t)W * C * dtype_size(dt);
    T->img.reserve((size_t)(hi - lo) * U.row_bytes);
    if (!h->copy_stream) Y3_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    U.dev = T->img.p; U.host = static_cast<const char*>(img); U.row_lo = lo; U.hi = hi; U.next_row = lo;
    // kernels of the previous call may still read T->img: order the first copy after the compute stream
    cudaEvent_t gate = take_event(h);
    Y3_CUDA(cudaEventRecord(gate, h->stream));
    Y3_CUDA(cudaStreamWaitEvent(h->copy_stream, gate, 0));
    U.ev.push_back(gate);
    return U;
}
// enqueue (copy stream) the rows up to row_end that are not on their way yet; returns the event to wait on
cudaEvent_t upload_rows_until(y3_context* h, BandUpload& U, int row_end) {
    if (!U.host) return nullptr;
    row_end = std::min(row_end, U.hi);
    if (row_end > U.next_row) {
        Y3_CUDA(cudaMemcpyAsync(const_cast<char*>(static_cast<const char*>(U.dev)) + (size_t)(U.next_row - U.row_lo) * U.row_bytes,
                                U.host + (size_t)U.next_row * U.row_bytes, (size_t)(row_end - U.next_row) * U.row_bytes,
                                cudaMemcpyHostToDevice, h->copy_stream));
        U.next_row = row_end;
    }
    cudaEvent_t e = take_event(h);
    Y3_CUDA(cudaEventRecord(e, h->copy_stream));
    U.ev.push_back(e);
    return e;
}
void release_upload(y3_context* h, BandUpload& U) {
    for (cudaEvent_t e : U.ev) h->event_pool.push_back(e);
    U.ev.clear();
}
const TileGeo* upload_geo(y3_context* h, Tiler* T, const std::vector<TileGeo>& geo) {
    T->geo.reserve(geo.size() * sizeof(TileGeo));
    Y3_CUDA(cudaMemcpyAsync(T->geo.p, geo.data(), geo.size() * sizeof(TileGeo), cudaMemcpyHostToDevice, h->stream));
    Y3_CUDA(cudaStreamSynchronize(h->stream));     // geo is a host temporary
    return T->geo.as<TileGeo>();
}
// Output of the tiled path.  Segmented pipeline: the rows are written by the NMS emit kernel straight into the caller's
// buffer when that lives on the device, else into T->acc (capacity = the caller's), and the row count stays on the
// device until finish_stitch().  Global-sort fallback: Tiler::stitch grows T->acc on the host side as before.
StitchCtx begin_stitch(y3_context* h, Tiler* T, PostProc* P, bool seg, double* preds, y3_mem mem, int64_t cap, const StitchArgs& S) {
    StitchCtx sc;
    sc.S = S;
    T->acc_rows = 0;
    if (!seg) return sc;
    sc.cap_rows = std::max<int64_t>(cap, 0);
    if (mem == Y3_MEM_DEVICE && preds) {
        sc.preds = preds;
    } else {
        T->acc.reserve((size_t)std::max<int64_t>(sc.cap_rows, 1) * 48);
        sc.preds = T->acc.as<double>();
    }
    P->begin_tiled();
    return sc;
}
// stats (optional): receives the control block of the segmented pipeline (candidate / kept totals)
void finish_stitch(y3_context* h, Tiler* T, PostProc* P, bool seg, double* preds, y3_mem mem, int64_t cap, int64_t* n_out, PostCtrl* stats) {
    if (seg) {
        const PostCtrl C = P->finish_tiled();
        if (stats) *stats = C;
        Y3_CHECK(!C.any_overflow, Y3_ERR_NOSPACE, "candidate list overflow in a tile batch (raise y3_config.max_candidates)");
        *n_out = (int64_t)C.acc_rows;
        Y3_CHECK(C.acc_rows <= cap, Y3_ERR_NOSPACE, "output capacity %lld < %lld boxes", (long long)cap, (long long)C.acc_rows);
        if (C.acc_rows && !(mem == Y3_MEM_DEVICE && preds)) {
            Y3_CHECK(preds, Y3_ERR_INVALID, "NULL output");
            from_device(h, preds, mem, T->acc.p, (size_t)C.acc_rows * 48);
        }
        Y3_CUDA(cudaStreamSynchronize(h->stream));
        return;
    }
    *n_out = T->acc_rows;
    Y3_CHECK(T->acc_rows <= cap, Y3_ERR_NOSPACE, "output capacity %lld < %lld boxes", (long long)cap, (long long)T->acc_rows);
    if (T->acc_rows) {
        Y3_CHECK(preds, Y3_ERR_INVALID, "NULL output");
        from_device(h, preds, mem, T->acc.p, (size_t)T->acc_rows * 48);
    }
    Y3_CUDA(cudaStreamSynchronize(h->stream));
}
}  // namespace

y3_status y3_tiles_normalized(y3_handle h, const void* img, y3_dtype dt, y3_mem img_mem, int64_t H, int64_t W, int32_t C,
                              int32_t th, int32_t tw, int32_t edge, int64_t first, int64_t count, float* out, y3_mem out_mem) {
    Y3_API_BEGIN(h)
    Y3_CHECK(img && out && H > 0 && W > 0 && C > 0, Y3_ERR_INVALID, "bad arguments");
    Tiler* T = tiler_of(h);
    std::vector<TileGeo> geo = plan_tiles(H, W, th, tw, edge, nullptr, nullptr);
    Y3_CHECK(first >= 0 && count >= 0 && first + count <= (int64_t)geo.size(), Y3_ERR_INVALID, "tile range [%lld,+%lld) outside 0..%zu",
             (long long)first, (long long)count, geo.size());
    if (count > 0) {
        long long row_lo = 0;
        const void* d_img = upload_band(h, T, img, dt, img_mem, W, C, geo, first, count, &row_lo);
        const TileGeo* d_geo = upload_geo(h, T, geo);
        const size_t bytes = (size_t)count * C * th * tw * 4;
        float* dst = out_mem == Y3_MEM_DEVICE ? out : (T->tiles.reserve(bytes), T->tiles.as<float>());
        T->sums.reserve((size_t)count * 16);
        launch_tile_norm(h, d_img, dt, row_lo, (int)W, C, d_geo + first, (int)count, th, tw, dst, nullptr, T->sums.as<double>());
        if (out_mem != Y3_MEM_DEVICE) from_device(h, out, out_mem, dst, bytes);
        Y3_CUDA(cudaStreamSynchronize(h->stream));
    }
    Y3_API_END(h)
}

y3_status y3_zscore(y3_handle h, const void* data, y3_dtype dt, y3_mem data_mem, int64_t n, float* out, y3_mem out_mem) {
    Y3_API_BEGIN(h)
    Y3_CHECK(data && out && n > 0 && n < (1ll << 31), Y3_ERR_INVALID, "bad arguments (1 <= n < 2^31)");
    Tiler* T = tiler_of(h);
    // one "tile" of th x tw = n elements, th = the largest power of two <= 4096 dividing n (rows spread over the CTAs)
    int th = 1;
    while (th < 4096 && n % (2 * th) == 0) th *= 2;
    const int tw = (int)(n / th);
    TileGeo g{};
    g.y0 = 0; g.y1 = th; g.x0 = 0; g.x1 = tw; g.pre_y = 0; g.pre_x = 0; g.rec_x = 0; g.rec_y = 0;
    const size_t in_bytes = (size_t)n * dtype_size(dt);
    const void* d_in = to_device(h, data, data_mem, in_bytes, T->img);
    T->geo.reserve(sizeof(TileGeo));
    Y3_CUDA(cudaMemcpyAsync(T->geo.p, &g, sizeof(TileGeo), cudaMemcpyHostToDevice, h->stream));
    Y3_CUDA(cudaStreamSynchronize(h->stream));     // g is a host temporary
    const size_t bytes = (size_t)n * 4;
    float* dst = out_mem == Y3_MEM_DEVICE ? out : (T->tiles.reserve(bytes), T->tiles.as<float>());
    T->sums.reserve(16);
    launch_tile_norm(h, d_in, dt, 0, tw, 1, T->geo.as<TileGeo>(), 1, th, tw, dst, nullptr, T->sums.as<double>());
    if (out_mem != Y3_MEM_DEVICE) from_device(h, out, out_mem, dst, bytes);
    Y3_CUDA(cudaStreamSynchronize(h->stream));
    Y3_API_END(h)
}

y3_status y3_tiles_raw(y3_handle h, const void* img, y3_dtype dt, y3_mem img_mem, int64_t H, int64_t W, int32_t C,
                       int32_t th, int32_t tw, int32_t edge, int64_t first, int64_t count, void* out, y3_mem out_mem) {
    Y3_API_BEGIN(h)
    Y3_CHECK(img && out && H > 0 && W > 0 && C > 0, Y3_ERR_INVALID, "bad arguments");
    Tiler* T = tiler_of(h);
    std::vector<TileGeo> geo = plan_tiles(H, W, th, tw, edge, nullptr, nullptr);
    Y3_CHECK(first >= 0 && count >= 0 && first + count <= (int64_t)geo.size(), Y3_ERR_INVALID, "tile range outside the plan");
    if (count > 0) {
        long long row_lo = 0;
        const void* d_img = upload_band(h, T, img, dt, img_mem, W, C, geo, first, count, &row_lo);
        const TileGeo* d_geo = upload_geo(h, T, geo);
        const size_t bytes = (size_t)count * C * th * tw * dtype_size(dt);
        void* dst = out_mem == Y3_MEM_DEVICE ? out : (T->tiles.reserve(bytes), T->tiles.p);
        launch_tile_raw(h, d_img, (int)dtype_size(dt), row_lo, (int)W, C, d_geo + first, (int)count, th, tw, dst);
        if (out_mem != Y3_MEM_DEVICE) from_device(h, out, out_mem, dst, bytes);
        Y3_CUDA(cudaStreamSynchronize(h->stream));
    }
    Y3_API_END(h)
}

y3_status y3_stitch_tiles(y3_handle h, const float* dets, y3_mem dets_mem, int64_t n_per_tile, int32_t nc, int64_t H, int64_t W,
                          int32_t th, int32_t tw, int32_t edge, int64_t first, int64_t count, float min_box, float iou_thr,
                          float score_thr, double* preds, y3_mem preds_mem, int64_t cap, int64_t* n_out) {
    Y3_API_BEGIN(h)
    Y3_CHECK(n_out && n_per_tile > 0 && nc >= 1 && (count == 0 || dets), Y3_ERR_INVALID, "bad arguments");
    Tiler* T = tiler_of(h);
    PostProc* P = post_of(h);
    std::vector<TileGeo> geo = plan_tiles(H, W, th, tw, edge, nullptr, nullptr);
    Y3_CHECK(first >= 0 && count >= 0 && first + count <= (int64_t)geo.size(), Y3_ERR_INVALID, "tile range outside the plan");
    T->acc_rows = 0;
    *n_out = 0;
    if (count > 0) {
        const TileGeo* d_geo = upload_geo(h, T, geo);
        const size_t per_tile = (size_t)n_per_tile * (5 + nc);
        StitchArgs S{H, W, th, tw, edge};
        // bounded batches so the candidate scratch stays small
        const int64_t step = std::max<int64_t>(1, std::min<int64_t>(count, (int64_t)(64ll << 20) / (int64_t)per_tile));
        const bool seg = PostProc::segmented_ok(dets_source(nullptr, n_per_tile, 1, nc, true, min_box, score_thr)) && !getenv("Y3_NMS_GLOBAL_SORT");
        StitchCtx sc = begin_stitch(h, T, P, seg, preds, preds_mem, cap, S);
        for (int64_t t0 = 0; t0 < count; t0 += step) {
            const int64_t nt = std::min(step, count - t0);
            const float* d = static_cast<const float*>(to_device(h, dets + (size_t)t0 * per_tile, dets_mem, (size_t)nt * per_tile * 4, T->dets));
            const CandSource src = dets_source(d, n_per_tile, (int)nt, nc, true, min_box, score_thr);
            sc.geo = d_geo + first + t0;
            if (!seg || !P->run_tiled(src, iou_thr, sc)) {
                NmsResult R = P->run(src, iou_thr);
                T->stitch(P, R, d_geo + first + t0, S);
            }
        }
        finish_stitch(h, T, P, seg, preds, preds_mem, cap, n_out, nullptr);
        flush_phases(h);
    }
    Y3_API_END(h)
}

namespace {
// inference_image_tiled for tiles [tile_first, tile_first + tile_count): shared by y3_infer_tiled and the sharded entry
void infer_tiled_impl(y3_handle h, const void* img, y3_dtype dt, y3_mem img_mem, int64_t H, int64_t W, int32_t C, int32_t th,
                      int32_t tw, int32_t edge, int64_t tile_first, int64_t tile_count, float min_box, float iou_thr,
                      float score_thr, double* preds, y3_mem preds_mem, int64_t cap, int64_t* n_out) {
    Net* net = net_of(h);
    Y3_CHECK(img && n_out, Y3_ERR_INVALID, "NULL pointer");
    Y3_CHECK(th == net->H && tw == net->W && C == net->C, Y3_ERR_INVALID,
             "tile %dx%dx%d does not match the network input %dx%dx%d", th, tw, C, net->H, net->W, net->C);
    Tiler* T = tiler_of(h);
    PostProc* P = post_of(h);
    std::vector<TileGeo> geo = plan_tiles(H, W, th, tw, edge, nullptr, nullptr);
    if (tile_count < 0) tile_count = (int64_t)geo.size() - tile_first;
    Y3_CHECK(tile_first >= 0 && tile_count >= 0 && tile_first + tile_count <= (int64_t)geo.size(), Y3_ERR_INVALID,
             "tile range [%lld,+%lld) outside 0..%zu", (long long)tile_first, (long long)tile_count, geo.size());
    y3_timings& Tm = h->timings;
    Tm.tiles = tile_count;
    Phase total(h, &Tm.ms_total, "y3:total");
    T->acc_rows = 0;
    *n_out = 0;
    if (tile_count > 0) {
        const TileGeo* d_geo = upload_geo(h, T, geo);
        BandUpload U = begin_band_upload(h, T, img, dt, img_mem, W, C, geo, tile_first, tile_count);
        const void* d_img = U.dev;
        const long long row_lo = U.row_lo;
        // Batch plan.  With the image in host memory the first batch is short (at most one row of tiles is worth waiting
        // for: the convolutions start after a small upload and the rest of the band streams in behind them).  The
        // remaining tiles are split EVENLY into as few batches as fit the network's capacity, but into at least three
        // batches per call (post-processing of batch k overlaps the convolutions of batch k+1; the last batch's
        // post-processing is exposed) unless that would make them smaller than 32 tiles.
        const std::vector<int> plan = batch_plan(tile_count, net->maxB, U.host != nullptr);
        auto rows_needed = [&](int64_t t0, int n) {           // last image row (exclusive) the batch [t0, t0 + n) reads
            int need = 0;
            for (int64_t t = 0; t < n; ++t) need = std::max(need, geo[tile_first + t0 + t].y1);
            return need;
        };
        cudaEvent_t ready = upload_rows_until(h, U, rows_needed(0, plan[0]));
        StitchArgs S{H, W, th, tw, edge};
        const int B = net->maxB;
        const bool seg = PostProc::segmented_ok(heads_source(net, 1, true, min_box, score_thr)) && !getenv("Y3_NMS_GLOBAL_SORT");
        StitchCtx sc = begin_stitch(h, T, P, seg, preds, preds_mem, cap, S);     // (main stream: ordered before every post-stream run)
        T->tiles.reserve((size_t)B * C * th * tw * 4);
        T->sums.reserve((size_t)B * 16);
        T->dbg_loop = 0.f;
        Phase loop_phase(h, &T->dbg_loop);
        // Software pipeline over tile batches: the conv stack of batch k+1 is enqueued on the main stream
        // BEFORE the post-processing of batch k (fused decode/threshold, sort, NMS, stitch - which contains
        // host synchronisations) runs on a second stream, reading the other set of head buffers.
        if (!h->post_stream) Y3_CUDA(cudaStreamCreateWithFlags(&h->post_stream, cudaStreamNonBlocking));
        static const bool overlap_post = getenv("Y3_NO_POST_OVERLAP") == nullptr;
        std::vector<cudaEvent_t> heads_ready;
        auto run_post = [&](int64_t t0, int nb, int set, cudaEvent_t ready_ev) {
            cudaStream_t main_stream = h->stream;
            if (overlap_post) {
                Y3_CUDA(cudaStreamWaitEvent(h->post_stream, ready_ev, 0));
                h->stream = h->post_stream;                    // every helper enqueues on ctx->stream
            }
            try {
                const CandSource src = heads_source(net, nb, true, min_box, score_thr, set);
                sc.geo = d_geo + tile_first + t0;
                bool done;
                { Phase p(h, &Tm.ms_nms, "y3:decode_nms_stitch"); done = seg && P->run_tiled(src, iou_thr, sc); p.stop(); }     // no host synchronisation
                if (!done) {
                    NmsResult R;
                    { Phase p(h, &Tm.ms_nms, "y3:decode_nms_stitch"); R = P->run(src, iou_thr); p.stop(); }
                    Tm.candidates += R.n_cand; Tm.kept += R.n_kept;
                    { Phase p(h, &Tm.ms_stitch, "y3:stitch"); T->stitch(P, R, d_geo + tile_first + t0, S); p.stop(); }
                }
            } catch (...) {
                h->stream = main_stream;
                throw;
            }
            h->stream = main_stream;
        };
        int64_t prev_t0 = -1; int prev_nb = 0, prev_set = 0; cudaEvent_t prev_ev = nullptr;
        int it = 0;
        int64_t t0 = 0;
        for (size_t bi = 0; bi < plan.size(); t0 += plan[bi], ++bi, ++it) {
            const int nb = plan[bi];
            const TileGeo* g = d_geo + tile_first + t0;
            const int set = it & 1;
            {   // wait for this batch's rows, then start the next batch's upload so it overlaps this batch's compute
                Phase p(h, &Tm.ms_h2d, "y3:h2d");
                if (ready) Y3_CUDA(cudaStreamWaitEvent(h->stream, ready, 0));
                p.stop();
                if (bi + 1 < plan.size()) ready = upload_rows_until(h, U, rows_needed(t0 + nb, plan[bi + 1]));
            }
            { Phase p(h, &Tm.ms_prep, "y3:tile_slice_zscore"); launch_tile_norm(h, d_img, dt, row_lo, (int)W, C, g, nb, th, tw, T->tiles.as<float>(), nullptr, T->sums.as<double>()); p.stop(); }
            { Phase p(h, &Tm.ms_conv, "y3:conv_stack"); net->forward(T->tiles.as<float>(), nb, set); p.stop(); }
            cudaEvent_t ev = take_event(h);
            Y3_CUDA(cudaEventRecord(ev, h->stream));
            heads_ready.push_back(ev);
            if (overlap_post) {
                if (prev_t0 >= 0) run_post(prev_t0, prev_nb, prev_set, prev_ev);     // batch k-1 while batch k computes
                prev_t0 = t0; prev_nb = nb; prev_set = set; prev_ev = ev;
            } else {
                run_post(t0, nb, set, ev);
            }
        }
        if (overlap_post && prev_t0 >= 0) run_post(prev_t0, prev_nb, prev_set, prev_ev);
        if (overlap_post) {                                     // the accumulated boxes live on the post stream
            cudaEvent_t done = take_event(h);
            Y3_CUDA(cudaEventRecord(done, h->post_stream));
            Y3_CUDA(cudaStreamWaitEvent(h->stream, done, 0));
            heads_ready.push_back(done);
        }
        for (cudaEvent_t e : heads_ready) h->event_pool.push_back(e);
        loop_phase.stop();
        release_upload(h, U);
        { Phase p(h, &Tm.ms_d2h, "y3:d2h");
          PostCtrl st{};
          finish_stitch(h, T, P, seg, preds, preds_mem, cap, n_out, &st);
          if (seg) { Tm.candidates = st.sum_cand; Tm.kept = st.sum_kept_nms; }
          p.stop(); }
    }
    total.stop();
}
}  // namespace

y3_status y3_infer_tiled(y3_handle h, const void* img, y3_dtype dt, y3_mem img_mem, int64_t H, int64_t W, int32_t C, int32_t th,
                         int32_t tw, int32_t edge, int64_t tile_first, int64_t tile_count, float min_box, float iou_thr,
                         float score_thr, double* preds, y3_mem preds_mem, int64_t cap, int64_t* n_out) {
    Y3_API_BEGIN(h)
    h->timings = y3_timings{};
    infer_tiled_impl(h, img, dt, img_mem, H, W, C, th, tw, edge, tile_first, tile_count, min_box, iou_thr, score_thr, preds, preds_mem,
                     cap, n_out);
    flush_phases(h);
    if (getenv("Y3_DEBUG_TIMING"))
        fprintf(stderr, "y3: total %.2f ms, stage sum %.2f ms\n", h->timings.ms_total,
                h->timings.ms_h2d + h->timings.ms_prep + h->timings.ms_conv + h->timings.ms_nms + h->timings.ms_stitch + h->timings.ms_d2h);
    h->timings.kernels_launched = h->kernels_launched;
    Y3_API_END(h)
}

y3_status y3_host_alloc(int64_t bytes, void** out) {
    if (!out || bytes <= 0) { g_create_error = "y3_host_alloc: bad arguments"; return Y3_ERR_INVALID; }
    *out = nullptr;
    const cudaError_t e = cudaHostAlloc(out, (size_t)bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();
        g_create_error = std::string("cudaHostAlloc failed: ") + cudaGetErrorString(e);
        return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? Y3_ERR_NODEVICE : Y3_ERR_CUDA;
    }
    return Y3_OK;
}

void y3_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

// ---------------------------------------------------------------------------------------- cross-seam stage
y3_status y3_cross_seam_nms(y3_handle h, const double* preds, y3_mem preds_mem, int64_t n, int32_t nc, int64_t H, int64_t W,
                            int32_t th, int32_t tw, int32_t edge, float iou_thr, double* out, y3_mem out_mem, int64_t cap, int64_t* n_out) {
    Y3_API_BEGIN(h)
    Y3_CHECK(n_out && n >= 0 && nc >= 1 && (n == 0 || preds), Y3_ERR_INVALID, "bad arguments");
    *n_out = 0;
    if (n > 0) {
        Tiler* T = tiler_of(h);
        const double* d = static_cast<const double*>(to_device(h, preds, preds_mem, (size_t)n * 48, T->dets));
        const StitchArgs S{H, W, th, tw, edge};
        const int64_t k = T->cross_seam(post_of(h), d, n, S, nc, iou_thr);
        *n_out = k;
        Y3_CHECK(k <= cap, Y3_ERR_NOSPACE, "output capacity %lld < %lld rows", (long long)cap, (long long)k);
        if (k) {
            Y3_CHECK(out, Y3_ERR_INVALID, "NULL output");
            from_device(h, out, out_mem, T->seam_out.p, (size_t)k * 48);
        }
        Y3_CUDA(cudaStreamSynchronize(h->stream));
        flush_phases(h);
    }
    Y3_API_END(h)
}

// ---------------------------------------------------------------------------------------- sharded path (NCCL)
y3_status y3_comm_unique_id(uint8_t* id) {
    if (!id) { g_create_error = "y3_comm_unique_id: NULL argument"; return Y3_ERR_INVALID; }
    try {
        comm_unique_id(id);
        return Y3_OK;
    } catch (const Error& e) {
        g_create_error = e.msg;
        return e.code;
    }
}

y3_status y3_comm_init(y3_handle h, int32_t rank, int32_t nranks, const uint8_t* id) {
    Y3_API_BEGIN(h)
    Y3_CHECK(nranks == 1 || id, Y3_ERR_INVALID, "NULL unique id");
    comm_init(h, rank, nranks, id);
    Y3_API_END(h)
}

int32_t y3_comm_size(y3_handle h) { return h ? comm_size(h) : 0; }

y3_status y3_infer_tiled_sharded(y3_handle h, const void* img, y3_dtype dt, y3_mem img_mem, int64_t H, int64_t W, int32_t C,
                                 int32_t th, int32_t tw, int32_t edge, float min_box, float iou_thr, float score_thr,
                                 int32_t cross_seam, double* preds, y3_mem preds_mem, int64_t cap, int64_t* n_out) {
    Y3_API_BEGIN(h)
    Y3_CHECK(n_out, Y3_ERR_INVALID, "NULL pointer");
    Y3_CHECK(h->comm, Y3_ERR_STATE, "y3_comm_init has not been called on this handle");
    Net* net = net_of(h);
    Tiler* T = tiler_of(h);
    const int rank = comm_rank(h), world = comm_size(h);
    Y3_CHECK(world <= 64, Y3_ERR_UNSUPPORTED, "more than 64 ranks");
    const int64_t n_tiles = (int64_t)plan_tiles(H, W, th, tw, edge, nullptr, nullptr).size();
    // contiguous shard of this rank: boundaries from the cumulative shares (equal at first, then following every rank's
    // measured throughput - identical on all ranks because they are derived from the same all-gathered numbers)
    const double* share = comm_shares(h);
    auto boundary = [&](int r) {
        double cum = 0;
        for (int k = 0; k < r; ++k) cum += share[k];
        return r >= world ? n_tiles : std::min<int64_t>(n_tiles, (int64_t)llround(cum * (double)n_tiles));
    };
    const int64_t first = boundary(rank);
    const int64_t count = boundary(rank + 1) - first;
    y3_timings& Tm = h->timings;
    Tm = y3_timings{};
    *n_out = 0;
    // 1. local shard -> device buffer.  An overflow must not keep this rank out of the collectives: it is carried
    //    through the count exchange and every rank reports Y3_ERR_NOSPACE together.
    const int64_t cap_local = std::max<int64_t>(cap, 1);
    T->shard_local.reserve((size_t)cap_local * 48);
    int64_t n_local = 0;
    bool local_overflow = false;
    cudaEvent_t t0 = take_event(h), t1 = take_event(h);
    Y3_CUDA(cudaEventRecord(t0, h->stream));
    try {
        infer_tiled_impl(h, img, dt, img_mem, H, W, C, th, tw, edge, first, count, min_box, iou_thr, score_thr,
                         T->shard_local.as<double>(), Y3_MEM_DEVICE, cap_local, &n_local);
    } catch (const Error& e) {
        if (e.code != Y3_ERR_NOSPACE) throw;
        local_overflow = true;
        for (auto& r_ : h->phase_log) { h->event_pool.push_back(r_.a); h->event_pool.push_back(r_.b); }
        h->phase_log.clear();
    }
    Y3_CUDA(cudaEventRecord(t1, h->stream));
    Y3_CUDA(cudaEventSynchronize(t1));
    float local_ms = 0.f;
    cudaEventElapsedTime(&local_ms, t0, t1);
    h->event_pool.push_back(t0); h->event_pool.push_back(t1);
    Tm.tiles = count;
    Phase comm_phase(h, &Tm.ms_comm, "y3:allgather");
    // 2. (row count, tiles, local microseconds) of every rank; a negative count flags an overflow
    T->shard_counts.reserve((size_t)(world + 1) * 3 * 8);
    h->pin_small.reserve(4096);
    long long* h_counts = h->pin_small.as<long long>();            // [world][3] gathered, then [3] local
    long long* h_mine = h_counts + 3 * 64;
    h_mine[0] = local_overflow ? -(long long)std::max<int64_t>(n_local, 1) : (long long)n_local;
    h_mine[1] = (long long)count;
    h_mine[2] = (long long)(local_ms * 1000.f);
    long long* d_counts = T->shard_counts.as<long long>();
    Y3_CUDA(cudaMemcpyAsync(d_counts + 3 * world, h_mine, 24, cudaMemcpyHostToDevice, h->stream));
    comm_all_gather_i64(h, d_counts + 3 * world, d_counts, 3);
    Y3_CUDA(cudaMemcpyAsync(h_counts, d_counts, (size_t)world * 24, cudaMemcpyDeviceToHost, h->stream));
    Y3_CUDA(cudaStreamSynchronize(h->stream));
    long long total = 0, max_c = 0;
    bool any_overflow = false;
    long long g_tiles[64], g_micros[64], g_rows[64];
    for (int r = 0; r < world; ++r) {
        const long long c = h_counts[3 * r];
        g_rows[r] = c; g_tiles[r] = h_counts[3 * r + 1]; g_micros[r] = h_counts[3 * r + 2];
        if (c < 0) any_overflow = true;
        total += c < 0 ? -c : c;
        max_c = std::max(max_c, c);
    }
    comm_update_shares(h, g_tiles, g_micros);                      // for the next call
    *n_out = total;
    Y3_CHECK(!any_overflow && total <= cap, Y3_ERR_NOSPACE, "output capacity %lld < %lld boxes", (long long)cap, total);
    if (total > 0) {
        // 3. records: one padded all-gather sized by the largest rank, then the ranks' rows are laid end to end
        //    (rank order = tile order, the reference's output order)
        T->shard_gather.reserve((size_t)world * max_c * 48);
        comm_all_gather_f64(h, T->shard_local.as<double>(), T->shard_gather.as<double>(), (size_t)max_c * 6);
        const bool to_caller = (preds_mem == Y3_MEM_DEVICE && preds && !cross_seam);
        if (!to_caller) T->acc.reserve((size_t)total * 48);
        double* dst = to_caller ? preds : T->acc.as<double>();
        long long off = 0;
        for (int r = 0; r < world; ++r) {
            if (g_rows[r] > 0)
                Y3_CUDA(cudaMemcpyAsync(dst + off * 6, T->shard_gather.as<double>() + (size_t)r * max_c * 6, (size_t)g_rows[r] * 48,
                                        cudaMemcpyDeviceToDevice, h->stream));
            off += g_rows[r];
        }
        int64_t n_final = total;
        const double* final_dev = dst;
        // 4. optional cross-seam NMS over the gathered boxes - identical on every rank
        if (cross_seam) {
            const StitchArgs S{H, W, th, tw, edge};
            n_final = T->cross_seam(post_of(h), dst, total, S, net->nc, iou_thr);
            final_dev = T->seam_out.as<double>();
            *n_out = n_final;
        }
        if (n_final > 0 && !to_caller) {
            Y3_CHECK(preds, Y3_ERR_INVALID, "NULL output");
            from_device(h, preds, preds_mem, final_dev, (size_t)n_final * 48);
        }
    }
    comm_phase.stop();
    Y3_CUDA(cudaStreamSynchronize(h->stream));
    flush_phases(h);
    Tm.kernels_launched = h->kernels_launched;
    Y3_API_END(h)
}

y3_status y3_get_timings(y3_handle h, y3_timings* out) {
    if (!h || !out) return Y3_ERR_INVALID;
    *out = h->timings;
    out->kernels_launched = h->kernels_launched;
    return Y3_OK;
}

y3_status y3_bench_forward(y3_handle h, int32_t batch, int32_t iters, float* ms_per_iter) {
    Y3_API_BEGIN(h)
    Net* net = net_of(h);
    Y3_CHECK(ms_per_iter && iters >= 1 && batch >= 1 && batch <= net->maxB, Y3_ERR_INVALID, "bad arguments");
    const size_t in_bytes = (size_t)batch * net->C * net->H * net->W * 4;
    h->stage_in.reserve(in_bytes);                 // whatever the last call left resident (or zeros)
    net->forward(h->stage_in.as<float>(), batch);  // warm-up
    float acc = 0.f;
    Phase p(h, &acc);
    for (int i = 0; i < iters; ++i) net->forward(h->stage_in.as<float>(), batch);
    p.stop();
    flush_phases(h);
    *ms_per_iter = acc / iters;
    Y3_API_END(h)
}

y3_status y3_profile_layers(y3_handle h, int32_t batch, int32_t iters, char* buf, int64_t cap) {
    Y3_API_BEGIN(h)
    Net* net = net_of(h);
    Y3_CHECK(buf && cap > 0 && iters >= 1 && batch >= 1 && batch <= net->maxB, Y3_ERR_INVALID, "bad arguments");
    const std::string rep = net->profile(batch, iters);
    const size_t n = std::min<size_t>(rep.size(), (size_t)cap - 1);
    memcpy(buf, rep.data(), n);
    buf[n] = 0;
    Y3_API_END(h)
}

y3_status y3_debug_layer_output(y3_handle h, const char* layer, int32_t batch, float* out, int64_t cap_floats, int32_t* dims) {
    Y3_API_BEGIN(h)
    Net* net = net_of(h);
    Y3_CHECK(layer && out && dims, Y3_ERR_INVALID, "NULL pointer");
    const Op* op = nullptr;
    for (const Op& o : net->ops) if (o.name == layer) { op = &o; break; }
    Y3_CHECK(op && op->kind != Op::DET, Y3_ERR_INVALID, "no such (non-detection) layer '%s'", layer);
    const TensorInfo& t = net->tensors[op->out.t];
    dims[0] = op->out.c; dims[1] = t.h; dims[2] = t.w;
    const int64_t n = (int64_t)batch * op->out.c * t.h * t.w;
    Y3_CHECK(n <= cap_floats, Y3_ERR_NOSPACE, "need %lld floats", (long long)n);
    h->stage_out.reserve((size_t)n * 4);
    slice_to_nchw(h, reinterpret_cast<const __nv_bfloat16*>(t.ptr), h->stage_out.as<float>(), batch, t.h, t.w, op->out.c, t.c, op->out.coff, t.f16);
    from_device(h, out, Y3_MEM_HOST, h->stage_out.p, (size_t)n * 4);
    Y3_CUDA(cudaStreamSynchronize(h->stream));
    Y3_API_END(h)
}

}  // extern "C"
