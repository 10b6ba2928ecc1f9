// net.cuh - static execution plan of the YOLOv3 graph (see net.cu).
#pragma once
#include "common.cuh"
#include "conv_tc.cuh"
#include "aux_kernels.cuh"

namespace y3 {

static constexpr int HEAD_PITCH_MAX = 256;   // fp32 channels per pixel of a stored head: A*(5+NC) padded to 32/64/128/256

struct TensorInfo {
    int h = 0, w = 0, c = 0;
    bool f16 = false;         // stored as fp16 instead of bf16 (tensors of the tail after the first upsample)
    int first = -1, last = -1;
    void* ptr = nullptr;
    size_t bytes = 0;
};

// channel slice [coff, coff+c) of NHWC tensor t
struct View {
    int t = -1, coff = 0, c = 0;
};

struct Op {
    enum Kind { STEM, CONV, DET, CONVT, UPCONV } kind = CONV;   // UPCONV = transposed conv + concat + 1x1 conv fused
    std::string name, bn, convt;
    int cin = 0, cout = 0, cout_pad = 0, k = 1, stride = 1;
    View in, in2, out, res;   // in2: the route half of a fused upsample+concat (UPCONV)
    int res_t = -1;
    int head = -1;
    bool flat = false;
    int pix_per_img = 0;
    unsigned have = 0;            // bit0 kernel, bit1 bias, bits2-5 gamma/beta/mean/var, bit6/7 convT kernel/bias
    DevBuf w, bias, scale, shift, bn_raw, raw_k, raw_b, raw_tk, raw_tb;
    std::vector<ConvLaunch> launches;
    Op() = default;
    Op(Op&& o) noexcept { *this = std::move(o); }
    Op& operator=(Op&& o) noexcept {
        kind = o.kind; name = std::move(o.name); bn = std::move(o.bn); convt = std::move(o.convt);
        cin = o.cin; cout = o.cout; cout_pad = o.cout_pad; k = o.k; stride = o.stride;
        in = o.in; in2 = o.in2; out = o.out; res = o.res; res_t = o.res_t; head = o.head; flat = o.flat;
        pix_per_img = o.pix_per_img; have = o.have; launches = std::move(o.launches);
        auto mv = [](DevBuf& a, DevBuf& b) { a.release(); a.p = b.p; a.cap = b.cap; b.p = nullptr; b.cap = 0; };
        mv(w, o.w); mv(bias, o.bias); mv(scale, o.scale); mv(shift, o.shift); mv(bn_raw, o.bn_raw);
        mv(raw_k, o.raw_k); mv(raw_b, o.raw_b); mv(raw_tk, o.raw_tk); mv(raw_tb, o.raw_tb);
        return *this;
    }
};

struct Net {
    y3_context* ctx;
    int H = 0, W = 0, C = 0, nc = 0, na = 0, maxB = 0, det_c = 0, head_pitch = 256;
    int gh[3], gw[3], row_start[3];
    int64_t rows_per_image = 0;
    double conv_flops_per_image = 0;
    size_t act_bytes = 0;
    std::vector<TensorInfo> tensors;
    std::vector<Op> ops;
    std::vector<void*> owned;
    float* head[3] = {nullptr, nullptr, nullptr};        // head set 0
    float* head_b[3] = {nullptr, nullptr, nullptr};      // head set 1 (post-processing of batch k overlaps conv of k+1)
    DevBuf boxes;                 // decoded [B, N, 5+NC] fp32
    DevBuf stage;
    bool loaded = false;
    bool fuse_stem_conv1 = false;   // ops[0] + ops[1] run as one kernel (conv_stem1.cu): 1-channel images only
    std::vector<float> stem_host;   // stem weights [9][32] | bias | scale | shift on the host (kernel parameters of the fused kernel)
    std::string missing = "all layers";
    int cur_batch = -1;
    int n_conv = 0, n_bn = 0, n_convt = 0;
    bool tail_f16 = false;          // tensors created from now on are fp16 (set after the first detection layer)

    explicit Net(y3_context* c) : ctx(c) {}
    ~Net();
    void build();
    void load(int n, const char* const* names, DLManagedTensor* const* tensors);
    void forward(const float* in_dev, int b, int head_set = 0);   // NCHW fp32 on the device -> heads[head_set]
    void forward_eager(const float* in_dev, int b, int head_set);
    // small batches replay a captured CUDA graph of the ~77 launches (the batch-1 forward is launch-bound)
    struct GraphEntry { const float* in; int b, head_set; cudaGraphExec_t exec; int launches; };
    std::vector<GraphEntry> graphs;
    void decode(int b);                         // heads -> boxes
    DecodeArgs decode_args(int b, int head_set = 0) const;   // for the fused decode+candidates path
    std::string profile(int b, int iters);      // per-layer CSV report

  private:
    int new_tensor(int h, int w, int c);
    View add_conv(View in, int cout, int k, int stride, int res_t = -1, View forced_out = View());
    View add_block(View x, int reps, View final_out);
    void add_yolo(View in, int f, View* route, View* out);
    void add_det(View in, int idx);
    void add_convt(View in, View out);
    View add_upconv(View x, View route, int cout);
    void add_yolo_up(View x, View route_in, int f, View* route, View* out);
    void make_launches(Op& op);
    void set_batch(Op& op, int b);
};

}  // namespace y3
