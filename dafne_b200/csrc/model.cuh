// Context of the dense forward: model description, packed weights, activation plan (see model.cu).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <functional>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "conv_tc.cuh"
#include "dafne_b200.h"
#include "postprocess.cuh"

namespace dafne {

struct ParamSlot {
    int64_t shape[4] = {1, 1, 1, 1};
    int64_t numel = 0;
    float* raw = nullptr;  // device fp32 copy in the reference's native layout
    bool loaded = false;
};

struct ConvPart {
    std::string prefix;  // state-dict prefix, e.g. "backbone.bottom_up.res2.0.conv1"
    int cout;
};

struct ConvLayer {
    std::vector<ConvPart> parts;  // >1 only for prediction convs fused along Cout
    int Cin = 0, Cout = 0, k = 1, stride = 1;
    bool bn = false;  // FrozenBN folded into scale/shift; otherwise conv bias -> shift
    __half* w = nullptr;
    float* scale = nullptr;
    float* shift = nullptr;
};

struct Act {
    __half* p = nullptr;
    int N = 0, H = 0, W = 0, C = 0;
    size_t off = 0, bytes = 0;
};

struct Arena {
    std::map<size_t, size_t> free_;  // offset -> size
    size_t top = 0, peak = 0;
    size_t alloc(size_t bytes);
    void release(size_t off, size_t bytes);
};

// What one planned launch is, for per-op timing: kind 0 = elementwise/other, 1 = tcgen05 conv.
struct OpInfo {
    int kind = 0;
    int block_n = 0, ksize = 0, stride = 0, Cin = 0, Cout = 0, Hout = 0, Wout = 0;
    double flops = 0;
    double bytes = 0;  // algorithmic HBM bytes (inputs + weights + outputs, each once)
    char name[48] = {0};
};

struct HeadOut {
    float* p = nullptr;
    int ld = 16;
    int H = 0, W = 0;
};

}  // namespace dafne

struct dafne_ctx {
    dafne_model_spec spec;
    int device = 0;
    int num_sms = 148;
    std::unordered_map<std::string, dafne::ParamSlot> params;
    std::vector<std::string> param_order;
    // layers
    std::vector<dafne::ConvLayer> convs;
    std::unordered_map<std::string, int> conv_index;  // first part prefix -> index in convs
    __half* stem_w = nullptr;                         // [64][7 ky][8 px][4 ch] fp16 (stem_tc.cu)
    float* stem_scale = nullptr;
    float* stem_shift = nullptr;
    float* scales_dev = nullptr;  // [L]
    bool finalized = false;
    // plan (valid after bind)
    int N = 0, H = 0, W = 0;
    uint8_t* ws = nullptr;
    __half* x0 = nullptr;  // preprocess output: zero-bordered NHWC4 fp16 canvas [N][H+6][W+8][4]
    size_t ws_bytes = 0;
    std::vector<std::function<int(cudaStream_t)>> ops;
    std::vector<dafne::OpInfo> op_info;  // parallel to ops (+ entry 0 = preprocess)
    bool profiling = false;
    std::vector<cudaEvent_t> prof_events;  // ops.size() + 2
    int64_t launches_per_forward = 0;
    double flops_per_forward = 0;
    long long* gn_sums_all = nullptr;
    size_t gn_sums_bytes = 0;
    int32_t* sizes_dev = nullptr;  // [N][4]
    void* images_dev = nullptr;    // staging for dafne_detect_host (slot 0 of the pipelined form)
    size_t images_dev_bytes = 0;
    float* dets_dev = nullptr;  // staging for dafne_detect_host
    int32_t* counts_dev = nullptr;
    int dets_capacity = 0;
    // dafne_detect_host_begin / _end: two batches in flight (slot 0 shares the buffers above)
    void* images_dev2 = nullptr;
    float* dets_dev2 = nullptr;
    int32_t* counts_dev2 = nullptr;
    cudaStream_t copy_stream = nullptr;  // H2D of the next batch
    cudaStream_t d2h_stream = nullptr;   // D2H of finished results (keeps the compute stream free of copies)
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_compute[2] = {nullptr, nullptr}, ev_result[2] = {nullptr, nullptr};
    bool slot_pending[2] = {false, false};
    bool slot_used[2] = {false, false};
    int slot_capacity[2] = {0, 0};  // rows per image of the slot's wire record (detections, then counts)
    unsigned slot_next = 0;
    dafne::HeadOut head_out[DAFNE_MAX_LEVELS][3];
    void* post_scratch = nullptr;
    size_t post_scratch_bytes = 0;
    // debugging / per-layer parity: keep every activation alive and addressable by name
    bool keep_activations = false;
    std::unordered_map<std::string, dafne::Act> named;
    // dafne_graph_capture / _launch: one whole step (dense forward + post-processing) as an instantiated CUDA graph
    cudaGraphExec_t graph_exec = nullptr;
    int64_t graph_launches = 0;  // kernel launches one replay stands for
    double graph_flops = 0;
    // counters
    int64_t stat_launches = 0;
    double stat_flops = 0;
};

namespace dafne {
int ctx_create(const dafne_model_spec* spec, int device, dafne_ctx** out);
void ctx_destroy(dafne_ctx* ctx);
int ctx_load_weights(dafne_ctx* ctx, int count, const char* const* names, const float* const* ptrs,
                     const int64_t* shapes, cudaStream_t s);
int ctx_finalize(dafne_ctx* ctx, cudaStream_t s);
int ctx_plan(dafne_ctx* ctx, int N, int H, int W, uint8_t* base, size_t bytes, size_t* needed);
int ctx_forward(dafne_ctx* ctx, const void* images, int dtype, const int32_t* image_sizes, cudaStream_t s);
int ctx_postprocess(dafne_ctx* ctx, const int32_t* image_sizes, const int32_t* output_sizes, int do_postprocess,
                    float* dets, int32_t* counts, int capacity, cudaStream_t s);
void fill_post_spec(const dafne_ctx* ctx, PostParams* p);
}  // namespace dafne
