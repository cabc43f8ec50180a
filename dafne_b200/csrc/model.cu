// Dense forward of the DAFNe detector as a static launch plan over caller-owned memory.
//
// Graph restated (semantics, not code) from:
//   detectron2 v0.5 ResNet (MSRA, STRIDE_IN_1X1, FrozenBN) + FPN(sum, no norm)   dafne/modeling/backbone/fpn.py:58-91
//   LastLevelP6P7: p6 = conv3x3s2(p5), p7 = conv3x3s2(relu(p6))                  dafne/modeling/backbone/fpn.py:16-37
//   DAFNeHead (center-to-corner, GN towers, CTR_ON_REG, corner tower on center)  dafne/modeling/dafne/dafne.py:287-348,388-414,462-471
// Weights arrive under their detectron2 state-dict names and are packed on device (fp16 [Cout][tap][Cin]; FrozenBN
// folded to per-channel scale/shift).
#include "model.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "elementwise.cuh"
#include "stem_tc.cuh"
#include "pair_tc.cuh"
#include "tail_tc.cuh"

namespace dafne {

#define CUDA_OK(expr)                                                                  \
    do {                                                                               \
        cudaError_t e__ = (expr);                                                      \
        if (e__ != cudaSuccess) {                                                      \
            set_error("%s: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return -1;                                                                 \
        }                                                                              \
    } while (0)

// ------------------------------------------------------------------------------------------------ arena
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

size_t Arena::alloc(size_t bytes) {
    bytes = align_up(bytes ? bytes : 1, 1024);
    // best fit among free blocks
    auto best = free_.end();
    for (auto it = free_.begin(); it != free_.end(); ++it)
        if (it->second >= bytes && (best == free_.end() || it->second < best->second)) best = it;
    if (best != free_.end()) {
        size_t off = best->first, sz = best->second;
        free_.erase(best);
        if (sz > bytes) free_[off + bytes] = sz - bytes;
        return off;
    }
    size_t off = top;
    top += bytes;
    if (top > peak) peak = top;
    return off;
}

void Arena::release(size_t off, size_t bytes) {
    bytes = align_up(bytes ? bytes : 1, 1024);
    auto it = free_.emplace(off, bytes).first;
    // coalesce with next
    auto nx = std::next(it);
    if (nx != free_.end() && it->first + it->second == nx->first) {
        it->second += nx->second;
        free_.erase(nx);
    }
    if (it != free_.begin()) {
        auto pv = std::prev(it);
        if (pv->first + pv->second == it->first) {
            pv->second += it->second;
            free_.erase(it);
            it = pv;
        }
    }
    if (it->first + it->second == top) {
        top = it->first;
        free_.erase(it);
    }
}

// ------------------------------------------------------------------------------------------------ model description
static int expect_param(dafne_ctx* c, const std::string& name, int64_t a, int64_t b = 1, int64_t d = 1, int64_t e = 1) {
    ParamSlot s;
    s.shape[0] = a;
    s.shape[1] = b;
    s.shape[2] = d;
    s.shape[3] = e;
    s.numel = a * b * d * e;
    CUDA_OK(cudaMalloc(&s.raw, s.numel * sizeof(float)));
    c->params.emplace(name, s);
    c->param_order.push_back(name);
    return 0;
}

static int add_conv(dafne_ctx* c, std::vector<ConvPart> parts, int Cin, int k, int stride, bool bn) {
    ConvLayer L;
    L.parts = parts;
    L.Cin = Cin;
    L.k = k;
    L.stride = stride;
    L.bn = bn;
    L.Cout = 0;
    for (auto& p : parts) {
        L.Cout += p.cout;
        if (expect_param(c, p.prefix + ".weight", p.cout, Cin, k, k)) return -1;
        if (bn) {
            for (const char* s : {".norm.weight", ".norm.bias", ".norm.running_mean", ".norm.running_var"})
                if (expect_param(c, p.prefix + s, p.cout)) return -1;
        } else {
            if (expect_param(c, p.prefix + ".bias", p.cout)) return -1;
        }
    }
    c->conv_index[parts[0].prefix] = static_cast<int>(c->convs.size());
    c->convs.push_back(L);
    return 0;
}

static const int* stage_blocks(int depth) {
    static const int r50[4] = {3, 4, 6, 3};
    static const int r101[4] = {3, 4, 23, 3};
    return depth == 50 ? r50 : (depth == 101 ? r101 : nullptr);
}

static const char* kTowers[3] = {"cls_tower", "center_tower", "corners_tower"};
static const std::string kBU = "backbone.bottom_up.";
static const std::string kHead = "proposal_generator.dafne_head.";

int ctx_create(const dafne_model_spec* spec, int device, dafne_ctx** out) {
    if (!spec || !out) {
        set_error("dafne_ctx_create: null argument");
        return -1;
    }
    if (!stage_blocks(spec->resnet_depth)) {
        set_error("dafne_ctx_create: resnet_depth %d unsupported (50 | 101)", spec->resnet_depth);
        return -1;
    }
    if (spec->num_levels != 5 || spec->num_classes < 1 || spec->num_classes > 32) {
        set_error("dafne_ctx_create: need num_levels == 5 and 1 <= num_classes <= 32 (got %d, %d)", spec->num_levels,
                  spec->num_classes);
        return -1;
    }
    CUDA_OK(cudaSetDevice(device));
    dafne_ctx* c = new dafne_ctx();
    c->spec = *spec;
    c->device = device;
    CUDA_OK(cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device));

    // stem
    if (expect_param(c, kBU + "stem.conv1.weight", 64, 3, 7, 7)) return -1;
    for (const char* s : {".norm.weight", ".norm.bias", ".norm.running_mean", ".norm.running_var"})
        if (expect_param(c, kBU + "stem.conv1" + s, 64)) return -1;
    // bottlenecks
    const int* nb = stage_blocks(spec->resnet_depth);
    int in_ch = 64, mid = 64, out_ch = 256;
    for (int s = 2; s <= 5; ++s) {
        for (int b = 0; b < nb[s - 2]; ++b) {
            const std::string pre = kBU + "res" + std::to_string(s) + "." + std::to_string(b);
            const int stride = (b == 0 && s > 2) ? 2 : 1;
            if (b == 0 && add_conv(c, {{pre + ".shortcut", out_ch}}, in_ch, 1, stride, true)) return -1;
            if (add_conv(c, {{pre + ".conv1", mid}}, in_ch, 1, stride, true)) return -1;  // STRIDE_IN_1X1
            if (add_conv(c, {{pre + ".conv2", mid}}, mid, 3, 1, true)) return -1;
            if (add_conv(c, {{pre + ".conv3", out_ch}}, mid, 1, 1, true)) return -1;
            in_ch = out_ch;
        }
        mid *= 2;
        out_ch *= 2;
    }
    // FPN
    const int feat_ch[3] = {512, 1024, 2048};
    for (int i = 0; i < 3; ++i) {
        if (add_conv(c, {{"backbone.fpn_lateral" + std::to_string(3 + i), 256}}, feat_ch[i], 1, 1, false)) return -1;
        if (add_conv(c, {{"backbone.fpn_output" + std::to_string(3 + i), 256}}, 256, 3, 1, false)) return -1;
    }
    if (add_conv(c, {{"backbone.top_block.p6", 256}}, 256, 3, 2, false)) return -1;
    if (add_conv(c, {{"backbone.top_block.p7", 256}}, 256, 3, 2, false)) return -1;
    // head
    for (int t = 0; t < 3; ++t)
        for (int i = 0; i < 4; ++i) {
            const std::string tw = kHead + kTowers[t] + ".";
            if (add_conv(c, {{tw + std::to_string(3 * i), 256}}, 256, 3, 1, false)) return -1;
            if (expect_param(c, tw + std::to_string(3 * i + 1) + ".weight", 256)) return -1;
            if (expect_param(c, tw + std::to_string(3 * i + 1) + ".bias", 256)) return -1;
        }
    if (add_conv(c, {{kHead + "cls_logits", spec->num_classes}}, 256, 3, 1, false)) return -1;
    if (add_conv(c, {{kHead + "ctrness", 1}, {kHead + "corners_pred", 8}}, 256, 3, 1, false)) return -1;
    if (add_conv(c, {{kHead + "center_pred", 2}}, 256, 3, 1, false)) return -1;
    for (int l = 0; l < 5; ++l)
        if (expect_param(c, kHead + "scales." + std::to_string(l) + ".scale", 1)) return -1;
    *out = c;
    return 0;
}

void ctx_destroy(dafne_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (auto& kv : c->params)
        if (kv.second.raw) cudaFree(kv.second.raw);
    for (auto& L : c->convs) {
        if (L.w) cudaFree(L.w);
        if (L.scale) cudaFree(L.scale);
        if (L.shift) cudaFree(L.shift);
    }
    if (c->stem_w) cudaFree(c->stem_w);
    if (c->stem_scale) cudaFree(c->stem_scale);
    if (c->stem_shift) cudaFree(c->stem_shift);
    if (c->scales_dev) cudaFree(c->scales_dev);
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
    for (int k = 0; k < 2; ++k) {
        if (c->ev_h2d[k]) cudaEventDestroy(c->ev_h2d[k]);
        if (c->ev_compute[k]) cudaEventDestroy(c->ev_compute[k]);
        if (c->ev_result[k]) cudaEventDestroy(c->ev_result[k]);
    }
    delete c;
}

int ctx_load_weights(dafne_ctx* c, int count, const char* const* names, const float* const* ptrs,
                     const int64_t* shapes, cudaStream_t s) {
    CUDA_OK(cudaSetDevice(c->device));
    for (int i = 0; i < count; ++i) {
        auto it = c->params.find(names[i]);
        if (it == c->params.end()) {
            set_error("dafne_load_weights: unknown tensor name '%s'", names[i]);
            return -1;
        }
        ParamSlot& p = it->second;
        const int64_t* sh = shapes + 4 * i;
        const int64_t numel = sh[0] * sh[1] * sh[2] * sh[3];
        bool same = numel == p.numel;
        for (int d = 0; d < 4 && same; ++d) same = sh[d] == p.shape[d];
        if (!same) {
            set_error("dafne_load_weights: '%s' has shape [%lld,%lld,%lld,%lld], expected [%lld,%lld,%lld,%lld]",
                      names[i], (long long)sh[0], (long long)sh[1], (long long)sh[2], (long long)sh[3],
                      (long long)p.shape[0], (long long)p.shape[1], (long long)p.shape[2], (long long)p.shape[3]);
            return -1;
        }
        CUDA_OK(cudaMemcpyAsync(p.raw, ptrs[i], numel * sizeof(float), cudaMemcpyDeviceToDevice, s));
        p.loaded = true;
    }
    c->finalized = false;
    return 0;
}

static float* raw_of(dafne_ctx* c, const std::string& name) { return c->params.at(name).raw; }

int ctx_finalize(dafne_ctx* c, cudaStream_t s) {
    CUDA_OK(cudaSetDevice(c->device));
    for (auto& name : c->param_order)
        if (!c->params.at(name).loaded) {
            set_error("dafne_weights_finalize: tensor '%s' was never loaded", name.c_str());
            return -1;
        }
    const float bn_eps = 1e-5f;  // detectron2 FrozenBatchNorm2d default
    if (!c->stem_w) {
        CUDA_OK(cudaMalloc(&c->stem_w, 64 * 224 * sizeof(__half)));
        CUDA_OK(cudaMalloc(&c->stem_scale, 64 * sizeof(float)));
        CUDA_OK(cudaMalloc(&c->stem_shift, 64 * sizeof(float)));
        CUDA_OK(cudaMalloc(&c->scales_dev, 8 * sizeof(float)));
    }
    const std::string st = kBU + "stem.conv1";
    if (launch_pack_stem_weight_tc(raw_of(c, st + ".weight"), c->stem_w, s)) return -1;
    if (launch_fold_bn(raw_of(c, st + ".norm.weight"), raw_of(c, st + ".norm.bias"),
                       raw_of(c, st + ".norm.running_mean"), raw_of(c, st + ".norm.running_var"), bn_eps, 64,
                       c->stem_scale, c->stem_shift, s))
        return -1;
    for (auto& L : c->convs) {
        const size_t kk = static_cast<size_t>(L.k) * L.k;
        if (!L.w) {
            CUDA_OK(cudaMalloc(&L.w, L.Cout * kk * L.Cin * sizeof(__half)));
            CUDA_OK(cudaMalloc(&L.shift, L.Cout * sizeof(float)));
            if (L.bn) CUDA_OK(cudaMalloc(&L.scale, L.Cout * sizeof(float)));
        }
        int row = 0;
        for (auto& p : L.parts) {
            if (launch_pack_conv_weight(raw_of(c, p.prefix + ".weight"), p.cout, L.Cin, L.k, L.w + row * kk * L.Cin, s))
                return -1;
            if (L.bn) {
                if (launch_fold_bn(raw_of(c, p.prefix + ".norm.weight"), raw_of(c, p.prefix + ".norm.bias"),
                                   raw_of(c, p.prefix + ".norm.running_mean"),
                                   raw_of(c, p.prefix + ".norm.running_var"), bn_eps, p.cout, L.scale + row,
                                   L.shift + row, s))
                    return -1;
            } else {
                CUDA_OK(cudaMemcpyAsync(L.shift + row, raw_of(c, p.prefix + ".bias"), p.cout * sizeof(float),
                                        cudaMemcpyDeviceToDevice, s));
            }
            row += p.cout;
        }
    }
    for (int l = 0; l < 5; ++l)
        CUDA_OK(cudaMemcpyAsync(c->scales_dev + l, raw_of(c, kHead + "scales." + std::to_string(l) + ".scale"),
                                sizeof(float), cudaMemcpyDeviceToDevice, s));
    c->finalized = true;
    return 0;
}

// ------------------------------------------------------------------------------------------------ plan
namespace {
constexpr size_t kMaxPlanProblems = 512;
constexpr size_t kMaxPairProblems = 80;  // 1x1 convolutions on the CTA-pair kernel: conv1 + conv3 of res4 / res5
constexpr size_t kMaxTailProblems = 40;  // bottleneck tails (conv3 + next conv1): 2 + 3 + 22 for ResNet-101
struct Builder {
    dafne_ctx* c;
    uint8_t* base;  // nullptr = dry run (size query)
    Arena arena;
    int64_t launches = 0;
    double flops = 0;
    bool failed = false;
    // convolutions are appended to `probs`; a launch covers the contiguous range of one group
    std::vector<ConvProblem> probs;
    ConvProblem* dev_probs = nullptr;
    bool grouping = false;
    size_t group_first = 0;
    int group_bn = 0, group_wgs = 0, group_res = -1, group_mode = -1, group_rs = -1, group_breg = -1;
    double group_flops = 0, group_bytes = 0;
    std::string group_name;
    OpInfo group_info;
    bool reverse_next = false;  // the next conv() walks its pixel tiles back to front (ConvParams::reverse_m)
    std::vector<TailProblem> tails;
    TailProblem* dev_tails = nullptr;
    std::vector<PairProblem> pairs;
    PairProblem* dev_pairs = nullptr;

    void begin_group(const std::string& name) {
        grouping = true;
        group_first = probs.size();
        group_bn = 0;
        group_wgs = 0;
        group_res = -1;
        group_mode = -1;
        group_rs = -1;
        group_breg = -1;
        group_flops = group_bytes = 0;
        group_name = name;
    }
    void end_group() {
        grouping = false;
        ++launches;
        if (!base || failed) return;
        const size_t first = group_first, cnt = probs.size() - first;
        if (cnt == 0 || cnt > static_cast<size_t>(kMaxConvProblems)) {
            set_error("plan: group '%s' has %zu problems", group_name.c_str(), cnt);
            failed = true;
            return;
        }
        int total = 0;
        for (size_t i = first; i < probs.size(); ++i) {
            probs[i].p.tile_begin = total;
            total += probs[i].p.total_tiles;
        }
        const ConvProblem* dp = dev_probs + first;
        const int bn = group_bn, wgs = group_wgs, res = group_res, mode = group_mode, rs = group_rs,
                  breg = group_breg, sms = c->num_sms, n = static_cast<int>(cnt);
        c->ops.push_back(
            [=](cudaStream_t s) { return conv_group_launch(dp, n, total, bn, wgs, mode, res, rs, breg, sms, s); });
        OpInfo o = group_info;
        snprintf(o.name, sizeof(o.name), "%s", group_name.size() > 46 ? group_name.substr(group_name.size() - 46).c_str()
                                                                      : group_name.c_str());
        o.kind = 1;
        o.flops = group_flops;
        o.bytes = group_bytes;
        o.block_n = bn;
        c->op_info.push_back(o);
    }

    Act new_act(int N, int H, int W, int C) {
        Act a;
        a.N = N;
        a.H = H;
        a.W = W;
        a.C = C;
        a.bytes = static_cast<size_t>(N) * H * W * C * sizeof(__half);
        a.off = arena.alloc(a.bytes);
        a.p = base ? reinterpret_cast<__half*>(base + a.off) : nullptr;
        return a;
    }
    void free_act(Act& a) {
        if (a.bytes && !c->keep_activations) arena.release(a.off, a.bytes);
        a.bytes = 0;
    }
    void info(const char* nm, int kind, double fl, double by, int bn = 0, int k = 0, int st = 0, int ci = 0,
              int co = 0, int ho = 0, int wo = 0) {
        if (!base) return;
        OpInfo o;
        snprintf(o.name, sizeof(o.name), "%s", nm);
        o.kind = kind;
        o.flops = fl;
        o.bytes = by;
        o.block_n = bn;
        o.ksize = k;
        o.stride = st;
        o.Cin = ci;
        o.Cout = co;
        o.Hout = ho;
        o.Wout = wo;
        c->op_info.push_back(o);
    }
    void name(const std::string& n, const Act& a) {
        if (base && c->keep_activations) c->named[n] = a;
    }
    template <typename T>
    T* persistent(size_t bytes) {
        size_t off = arena.alloc(bytes);
        return base ? reinterpret_cast<T*>(base + off) : nullptr;
    }
    const ConvLayer& layer(const std::string& prefix) { return c->convs[c->conv_index.at(prefix)]; }

    // Bottleneck tail (tail_tc.cu): out = relu(bn3(conv3(in)) + sc) and mid = relu(bn1'(conv1'(out))) in one launch
    void tail(const ConvLayer& L3, const ConvLayer& L1, const Act& in, const Act& sc, Act* out, Act* mid,
              const std::string& nm) {
        *out = new_act(in.N, in.H, in.W, L3.Cout);
        *mid = new_act(in.N, in.H, in.W, L1.Cout);
        ++launches;
        const double fl = 2.0 * in.N * in.H * in.W * ((double)L3.Cout * L3.Cin + (double)L1.Cout * L1.Cin);
        flops += fl;
        if (!base || failed) return;
        if (in.C != L3.Cin || sc.C != L3.Cout || L1.Cin != L3.Cout || tails.size() >= kMaxTailProblems) {
            set_error("plan: bottleneck tail '%s' has inconsistent shapes", nm.c_str());
            failed = true;
            return;
        }
        TailDesc d;
        d.in = in.p;
        d.N = in.N;
        d.H = in.H;
        d.W = in.W;
        d.K1 = L3.Cin;
        d.w3 = L3.w;
        d.N1 = L3.Cout;
        d.scale1 = L3.scale;
        d.shift1 = L3.shift;
        d.residual = sc.p;
        d.out = out->p;
        d.w1 = L1.w;
        d.N2 = L1.Cout;
        d.scale2 = L1.scale;
        d.shift2 = L1.shift;
        d.mid = mid->p;
        TailPlan plan;
        if (tail_plan_build(d, &plan, c->num_sms)) {
            failed = true;
            return;
        }
        const TailProblem* dp = dev_tails + tails.size();
        tails.push_back(plan.prob);
        c->ops.push_back([dp, plan](cudaStream_t s) { return tail_plan_launch(dp, plan, s); });
        const double px = (double)in.N * in.H * in.W;
        info(nm.size() > 46 ? nm.substr(nm.size() - 46).c_str() : nm.c_str(), 1, fl,
             px * 2.0 * (L3.Cin + 2.0 * L3.Cout + L1.Cout) + 2.0 * ((double)L3.Cout * L3.Cin + (double)L1.Cout * L1.Cin),
             128, 1, 1, L3.Cin, L3.Cout, in.H, in.W);
    }

    // 1x1 / stride-1 convolution + FrozenBN (+ shortcut) (+ ReLU) on the CTA-pair kernel (pair_tc.cu)
    Act pair(const ConvLayer& L, const Act& in, bool relu, const Act* residual, const std::string& nm) {
        Act out = new_act(in.N, in.H, in.W, L.Cout);
        ++launches;
        const double px = (double)in.N * in.H * in.W;
        const double fl = 2.0 * px * (double)L.Cout * L.Cin;
        flops += fl;
        const bool rev = reverse_next;
        reverse_next = false;
        if (!base || failed) return out;
        if (in.C != L.Cin || (residual && (residual->C != L.Cout || residual->H != in.H || residual->W != in.W)) ||
            pairs.size() >= kMaxPairProblems) {
            set_error("plan: pair convolution '%s' has inconsistent shapes", nm.c_str());
            failed = true;
            return out;
        }
        PairDesc d;
        d.in = in.p;
        d.M = (long long)in.N * in.H * in.W;
        d.K = L.Cin;
        d.w = L.w;
        d.N = L.Cout;
        d.scale = L.scale;
        d.shift = L.shift;
        d.relu = relu ? 1 : 0;
        d.residual = residual ? residual->p : nullptr;
        d.out = out.p;
        d.reverse_m = rev ? 1 : 0;
        PairPlan plan;
        if (pair_plan_build(d, &plan, c->num_sms)) {
            failed = true;
            return out;
        }
        const PairProblem* dp = dev_pairs + pairs.size();
        pairs.push_back(plan.prob);
        c->ops.push_back([dp, plan](cudaStream_t s) { return pair_plan_launch(dp, plan, s); });
        info(nm.size() > 46 ? nm.substr(nm.size() - 46).c_str() : nm.c_str(), 1, fl, plan.bytes, 256, 1, 1, L.Cin, L.Cout,
             in.H, in.W);
        return out;
    }
    static bool pair_ok(const ConvLayer& L) {
        return L.k == 1 && L.stride == 1 && L.scale != nullptr && L.shift != nullptr && pair_supported(L.Cin, L.Cout);
    }

    // out = epilogue(conv(in)); allocates the fp16 output unless out_f32 is given
    // in_gn_sums / in_gamma / in_beta: `in` is a RAW tower output; its GroupNorm + ReLU is applied while it is loaded
    Act conv(const ConvLayer& L, const Act& in, bool relu, const Act* residual = nullptr, int res_shift = 0,
             long long* gn_sums = nullptr, float* out_f32 = nullptr, int out_ld = 0,
             const long long* in_gn_sums = nullptr, const float* in_gamma = nullptr, const float* in_beta = nullptr) {
        ConvDesc d;
        d.in = in.p;
        d.N = in.N;
        d.Hin = in.H;
        d.Win = in.W;
        d.Cin = L.Cin;
        d.w = L.w;
        d.Cout = L.Cout;
        d.ksize = L.k;
        d.stride = L.stride;
        conv_out_dims(d);
        Act out;
        if (out_f32) {
            out.N = in.N;
            out.H = d.Hout;
            out.W = d.Wout;
            out.C = L.Cout;
        } else {
            out = new_act(in.N, d.Hout, d.Wout, L.Cout);
        }
        d.out = out.p;
        d.out_f32 = out_f32;
        d.out_ld = out_ld;
        d.scale = L.scale;
        d.shift = L.shift;
        d.relu = relu ? 1 : 0;
        if (residual) {
            d.residual = residual->p;
            d.res_H = residual->H;
            d.res_W = residual->W;
            d.res_shift = res_shift;
        }
        d.gn_sums = gn_sums;
        d.in_gn_sums = in_gn_sums;
        d.in_gamma = in_gamma;
        d.in_beta = in_beta;
        d.reverse_m = reverse_next ? 1 : 0;
        reverse_next = false;
        const bool single = !grouping;
        if (single) begin_group(L.parts[0].prefix);
        flops += 2.0 * in.N * d.Hout * d.Wout * (double)L.Cout * L.k * L.k * L.Cin;
        if (base) {
            if (in.C != L.Cin) {
                set_error("plan: conv '%s' expects Cin=%d, got %d", L.parts[0].prefix.c_str(), L.Cin, in.C);
                failed = true;
                return out;
            }
            ConvPlan plan;
            if (conv_plan_build(d, &plan, c->num_sms)) {
                failed = true;
                return out;
            }
            if (group_bn != 0 && group_bn != plan.block_n) {
                set_error("plan: group '%s' mixes tile widths %d and %d", group_name.c_str(), group_bn, plan.block_n);
                failed = true;
                return out;
            }
            if (probs.size() >= kMaxPlanProblems) {
                set_error("plan: more than %d convolutions", kMaxPlanProblems);
                failed = true;
                return out;
            }
            if ((group_res >= 0 && group_res != plan.res_tma) || (group_mode >= 0 && group_mode != plan.mode) ||
                (group_rs >= 0 && group_rs != plan.row_shared) || (group_breg >= 0 && group_breg != plan.breg_bytes)) {
                set_error("plan: group '%s' mixes TMA-residual and other convolutions", group_name.c_str());
                failed = true;
                return out;
            }
            group_res = plan.res_tma;
            group_mode = plan.mode;
            group_rs = plan.row_shared;
            group_breg = plan.breg_bytes;
            group_bn = plan.block_n;
            group_wgs = group_wgs == 0 ? plan.epi_wgs : (group_wgs < plan.epi_wgs ? group_wgs : plan.epi_wgs);
            probs.push_back(plan.prob);
            const double px_in = (double)in.N * in.H * in.W, px_out = (double)in.N * d.Hout * d.Wout;
            group_bytes += px_in * L.Cin * 2 + (double)L.Cout * L.k * L.k * L.Cin * 2 +
                           px_out * (out_f32 ? out_ld * 4.0 : L.Cout * 2.0) +
                           (residual ? (double)residual->N * residual->H * residual->W * L.Cout * 2 : 0.0);
            group_flops += plan.flops;
            if (probs.size() - group_first == 1) {
                group_info = OpInfo();
                group_info.ksize = L.k;
                group_info.stride = L.stride;
                group_info.Cin = L.Cin;
                group_info.Cout = L.Cout;
                group_info.Hout = d.Hout;
                group_info.Wout = d.Wout;
            }
        }
        if (single) end_group();
        return out;
    }
};
}  // namespace

int ctx_plan(dafne_ctx* c, int N, int H, int W, uint8_t* base, size_t bytes, size_t* needed) {
    if (N < 1 || H < 32 || W < 32 || H % 32 || W % 32) {
        set_error("plan: need N >= 1 and H, W positive multiples of 32 (got N=%d H=%d W=%d)", N, H, W);
        return -1;
    }
    if (base && !c->finalized) {
        set_error("plan: call dafne_weights_finalize before dafne_bind_workspace");
        return -1;
    }
    CUDA_OK(cudaSetDevice(c->device));
    if (base && (c->slot_pending[0] || c->slot_pending[1])) {
        set_error("bind_workspace: a pipelined batch is still in flight (call dafne_detect_host_end first)");
        return -1;
    }
    if (base && c->copy_stream) {
        // copies of an earlier plan may still be running on the context's private streams: they must not outlive the
        // workspace they read / write
        CUDA_OK(cudaStreamSynchronize(c->copy_stream));
        CUDA_OK(cudaStreamSynchronize(c->d2h_stream));
    }
    Builder B;
    B.c = c;
    B.base = base;
    if (base) {
        if (c->graph_exec) {  // a captured step refers to the old plan's buffers
            cudaGraphExecDestroy(c->graph_exec);
            c->graph_exec = nullptr;
        }
        c->ops.clear();
        c->named.clear();
        c->op_info.clear();
        B.info("preprocess", 0, 0, (double)N * H * W * 3 + (double)N * (H + 6) * (W + 8) * 8);
    }
    const dafne_model_spec& sp = c->spec;

    // ---- persistent buffers
    int lvH[5], lvW[5];
    lvH[0] = H / 8;
    lvW[0] = W / 8;
    lvH[1] = H / 16;
    lvW[1] = W / 16;
    lvH[2] = H / 32;
    lvW[2] = W / 32;
    for (int l = 3; l < 5; ++l) {
        lvH[l] = (lvH[l - 1] - 1) / 2 + 1;
        lvW[l] = (lvW[l - 1] - 1) / 2 + 1;
    }
    int32_t* sizes_dev = B.persistent<int32_t>(static_cast<size_t>(N) * 4 * sizeof(int32_t));
    B.dev_probs = B.persistent<ConvProblem>(kMaxPlanProblems * sizeof(ConvProblem));
    B.dev_tails = B.persistent<TailProblem>(kMaxTailProblems * sizeof(TailProblem));
    B.dev_pairs = B.persistent<PairProblem>(kMaxPairProblems * sizeof(PairProblem));
    const size_t sums_per = static_cast<size_t>(N) * 32 * 2 * sizeof(long long);
    const size_t sums_bytes = sums_per * 3 * 4 * 5;
    long long* sums_all = B.persistent<long long>(sums_bytes);
    HeadOut ho[5][3];
    const int ld_logits = sp.num_classes <= 16 ? 16 : 32;
    for (int l = 0; l < 5; ++l)
        for (int k = 0; k < 3; ++k) {
            ho[l][k].ld = k == 0 ? ld_logits : 16;
            ho[l][k].H = lvH[l];
            ho[l][k].W = lvW[l];
            ho[l][k].p = B.persistent<float>(static_cast<size_t>(N) * lvH[l] * lvW[l] * ho[l][k].ld * sizeof(float));
        }
    int level_hw[10];
    for (int l = 0; l < 5; ++l) {
        level_hw[2 * l] = lvH[l];
        level_hw[2 * l + 1] = lvW[l];
    }
    const size_t post_bytes = postprocess_scratch_bytes(N, 5, level_hw, sp.num_classes, sp.pre_nms_topk);
    void* post_scratch = B.persistent<uint8_t>(post_bytes);
    const size_t img_bytes = static_cast<size_t>(N) * 3 * H * W * sizeof(float);
    void* images_dev = B.persistent<uint8_t>(img_bytes);
    const int det_cap = 2048;
    // + N int32: the pipelined host call keeps a slot's counts right behind its detections (dafne_host_slot_wire)
    const size_t dets_bytes = static_cast<size_t>(N) * det_cap * DAFNE_DET_STRIDE * sizeof(float) + static_cast<size_t>(N) * 4;
    float* dets_dev = B.persistent<float>(dets_bytes);
    int32_t* counts_dev = B.persistent<int32_t>(static_cast<size_t>(N) * sizeof(int32_t));
    void* images_dev2 = B.persistent<uint8_t>(img_bytes);
    float* dets_dev2 = B.persistent<float>(dets_bytes);
    int32_t* counts_dev2 = B.persistent<int32_t>(static_cast<size_t>(N) * sizeof(int32_t));

    // ---- stem (tensor cores; the preprocess kernel writes the zero-bordered NHWC4 canvas it reads)
    Act x0;
    x0.N = N;
    x0.H = H + 6;
    x0.W = W + 8;
    x0.C = 4;
    x0.bytes = static_cast<size_t>(N) * (H + 6) * (W + 8) * 4 * sizeof(__half);
    x0.off = B.arena.alloc(x0.bytes);
    x0.p = base ? reinterpret_cast<__half*>(base + x0.off) : nullptr;
    __half* x0p_saved = x0.p;
    // stem conv + FrozenBN + ReLU + max-pool in one launch: the H/2 x W/2 conv output never reaches HBM
    Act x = B.new_act(N, (H / 2 - 1) / 2 + 1, (W / 2 - 1) / 2 + 1, 64);
    B.launches += 2;
    B.flops += 2.0 * N * (H / 2) * (W / 2) * 64.0 * 49 * 3;
    if (base) {
        // op 0 (preprocess) is issued by ctx_forward because it takes the per-call image pointer
        StemPlan sp_;
        if (stem_plan_build(x0.p, N, H, W, c->stem_w, c->stem_scale, c->stem_shift, x.p, &sp_, c->num_sms)) return -1;
        c->ops.push_back([sp_](cudaStream_t s) { return stem_plan_launch(sp_, s); });
        B.info("stem+pool", 2, 2.0 * N * (H / 2) * (W / 2) * 64.0 * 147, (double)x0.bytes + (double)x.bytes);
    }
    // x0 is needed until the stem ran; releasing at plan time is safe because ops execute in plan order
    B.name("pool", x);
    B.free_act(x0);

    // ---- bottlenecks
    const int* nb = stage_blocks(sp.resnet_depth);
    Act feats[3];
    // Which stages run conv3 + the next block's conv1 as one fused tail launch: bit (s - 2) of DAFNE_CONV_TAIL
    // (default 3 = res2 and res3; 0 = every 1x1 is its own launch; 7 adds res4). Measured at R101 32 x 1024^2 with the
    // final kernels: the fused tail moves 28 % fewer bytes but at 3.6 TB/s where the two plain launches run at 5.7-6.2 --
    // res2 738 vs 421 + 219 us, res3 378 vs 218 + 116 us per block pair -- and the power-capped STEP is the same either way
    // (three interleaved same-box rounds: 1 098 / 1 097 / 1 077 images/s with 0, 1 094 / 1 090 / 1 088 with 3): fewer bytes
    // buy back as clock what the longer launch costs. res4 (257 vs 134 + 79 us) is slower in both views.
    static const int tail_mask = getenv("DAFNE_CONV_TAIL") ? atoi(getenv("DAFNE_CONV_TAIL")) : 3;
    // Which stages run their 1x1 / stride-1 convolutions on the CTA-pair kernel (pair_tc.cu, cta_group::2): bit (s - 2) of
    // DAFNE_CONV_PAIR for conv1, bit (s - 2 + 4) for conv3 (0 = none). Default 0x8C = conv1 of res4 / res5 and conv3 of
    // res5 (per launch at R101 32 x 1024^2: 79 -> 70 us, 75 -> 64 us, 101 -> 84 us); conv3 of res4 is SLOWER on the pair
    // kernel (134 -> 146 us: K = 256 gives a 256 x 256 tile only four k-blocks of tensor work per 192 KB of HBM traffic, and
    // three 32 KB residual slots prefetch less far ahead than conv_tc.cu's ring of 16 KB slots).
    static const int pair_mask = getenv("DAFNE_CONV_PAIR") ? atoi(getenv("DAFNE_CONV_PAIR")) : 0x8C;
    for (int s = 2; s <= 5; ++s) {
        Act next_a;  // conv1 output of the NEXT block when the previous block's tail has already produced it
        bool have_a = false;
        for (int b = 0; b < nb[s - 2]; ++b) {
            const std::string pre = kBU + "res" + std::to_string(s) + "." + std::to_string(b);
            Act sc = x;
            bool own_sc = false;
            if (b == 0) {
                sc = B.conv(B.layer(pre + ".shortcut"), x, false);
                own_sc = true;
            }
            Act a;
            if (have_a) {
                a = next_a;
                have_a = false;
            } else {
                // conv1 reads the block input the previous block's conv3 has just written (268 MB at 32 x 1024^2 in
                // res4, twice the L2): back to front, the part L2 still holds comes first. DAFNE_CONV_REVERSE=0: A/B.
                static const bool reverse_on = !(getenv("DAFNE_CONV_REVERSE") && atoi(getenv("DAFNE_CONV_REVERSE")) == 0);
                B.reverse_next = reverse_on && b > 0;
                const ConvLayer& L1 = B.layer(pre + ".conv1");
                a = (((pair_mask >> (s - 2)) & 1) && Builder::pair_ok(L1)) ? B.pair(L1, x, true, nullptr, pre + ".conv1")
                                                                             : B.conv(L1, x, true);
            }
            Act m = B.conv(B.layer(pre + ".conv2"), a, true);
            B.free_act(a);
            Act o;
            const ConvLayer& L3 = B.layer(pre + ".conv3");
            bool fused = false;
            if (((tail_mask >> (s - 2)) & 1) && b + 1 < nb[s - 2]) {
                const ConvLayer& L1n = B.layer(kBU + "res" + std::to_string(s) + "." + std::to_string(b + 1) + ".conv1");
                if (L1n.stride == 1 && L1n.k == 1 && tail_supported(L3.Cin, L3.Cout, L1n.Cout)) {
                    // conv3 of this block + conv1 of the next one: the block output is written once, never read back
                    B.tail(L3, L1n, m, sc, &o, &next_a, pre + ".conv3+next.conv1");
                    have_a = true;
                    fused = true;
                }
            }
            if (!fused)
                o = (((pair_mask >> (s - 2 + 4)) & 1) && Builder::pair_ok(L3)) ? B.pair(L3, m, true, &sc, pre + ".conv3")
                                                                                 : B.conv(L3, m, true, &sc, 0);
            B.free_act(m);
            if (own_sc) B.free_act(sc);
            B.free_act(x);  // block input (== sc for b > 0)
            x = o;
            B.name("res" + std::to_string(s) + "." + std::to_string(b), o);
            if (B.failed) return -1;
        }
        if (s >= 3) {
            feats[s - 3] = x;
            // keep the stage output alive for the FPN lateral: hand the next stage a non-owning alias
            if (s < 5) x.bytes = 0;
        }
    }
    feats[2] = x;

    // ---- FPN (top-down), P6, P7. The lateral convs form a chain (each adds the upsampled previous one); the three
    // output convs only depend on their own lateral sum, so they run as ONE grouped launch afterwards.
    Act inner[3], P[5];
    for (int i = 2; i >= 0; --i) {
        const std::string lat = "backbone.fpn_lateral" + std::to_string(3 + i);
        inner[i] = (i == 2) ? B.conv(B.layer(lat), feats[i], false) : B.conv(B.layer(lat), feats[i], false, &inner[i + 1], 1);
        B.free_act(feats[i]);
        if (B.failed) return -1;
    }
    B.begin_group("backbone.fpn_output3+4+5");
    for (int i = 0; i < 3; ++i) P[i] = B.conv(B.layer("backbone.fpn_output" + std::to_string(3 + i)), inner[i], false);
    B.end_group();
    if (B.failed) return -1;
    for (int i = 0; i < 3; ++i) B.free_act(inner[i]);
    for (int i = 0; i < 3; ++i) B.name("p" + std::to_string(3 + i), P[i]);
    P[3] = B.conv(B.layer("backbone.top_block.p6"), P[2], false);
    Act p6r = B.new_act(P[3].N, P[3].H, P[3].W, 256);
    B.launches += 1;
    if (base) {
        const __half* src = P[3].p;
        __half* dst = p6r.p;
        const size_t n8 = static_cast<size_t>(P[3].N) * P[3].H * P[3].W * 256 / 8;
        c->ops.push_back([=](cudaStream_t s) { return launch_relu_copy(src, dst, n8, s); });
        B.info("relu_p6", 0, 0, (double)n8 * 32);
    }
    P[4] = B.conv(B.layer("backbone.top_block.p7"), p6r, false);
    B.free_act(p6r);
    B.name("p6", P[3]);
    B.name("p7", P[4]);
    if (B.failed) return -1;

    // ---- head: three 4-conv GN towers + prediction convs; the weights are shared by the five levels, so every
    // tower layer is ONE grouped launch over all levels (and over the two towers that run side by side):
    //   cls_tower(f), center_tower(f), corners_tower(center_tower(f))                        (dafne.py:362-404)
    for (int l = 0; l < 5; ++l)
        if (P[l].H != lvH[l] || P[l].W != lvW[l]) {
            set_error("plan: level %d is %dx%d, expected %dx%d", l, P[l].H, P[l].W, lvH[l], lvW[l]);
            return -1;
        }
    auto sums_of = [&](int t, int i, int l) -> long long* {
        return sums_all ? sums_all + ((static_cast<size_t>(t) * 4 + i) * 5 + l) * (sums_per / sizeof(long long))
                        : reinterpret_cast<long long*>(1);
    };
    auto gn_group = [&](const std::vector<std::pair<int, Act*>>& items, int layer) {
        // items: (tower index, activation) -> normalise in place
        B.launches += 1;
        if (!base) return;
        std::vector<GnProblem> gp;
        double bytes = 0;
        int k = 0;
        for (auto& it : items) {
            const std::string tw = kHead + kTowers[it.first] + ".";
            GnProblem g;
            g.x = it.second->p;
            g.sums = sums_of(it.first, layer, k % 5);
            g.gamma = raw_of(c, tw + std::to_string(3 * layer + 1) + ".weight");
            g.beta = raw_of(c, tw + std::to_string(3 * layer + 1) + ".bias");
            g.N = N;
            g.HW = it.second->H * it.second->W;
            gp.push_back(g);
            bytes += (double)N * g.HW * 256 * 4;
            ++k;
        }
        c->ops.push_back([gp](cudaStream_t s) { return launch_gn_relu_group(gp.data(), (int)gp.size(), 1e-5f, s); });
        B.info("gn_relu", 0, 0, bytes);
    };
    // GroupNorm + ReLU between the tower layers is applied by the CONSUMING convolution while it loads its input
    // (conv_tc mode 5): layers 0-2 of a tower store their raw output and statistics only; the explicit in-place pass
    // remains for layer 3, whose consumers are the narrow prediction convolutions and corners_tower.0.
    // DAFNE_CONV_GNFUSE=0 (or DAFNE_CONV_HALO256=0): every layer is followed by the separate pass again.
    const bool gn_fuse = !(getenv("DAFNE_CONV_GNFUSE") && atoi(getenv("DAFNE_CONV_GNFUSE")) == 0) &&
                         !(getenv("DAFNE_CONV_HALO256") && atoi(getenv("DAFNE_CONV_HALO256")) == 0);
    auto gn_w = [&](int t, int layer) { return raw_of(c, kHead + kTowers[t] + "." + std::to_string(3 * layer + 1) + ".weight"); };
    auto gn_b = [&](int t, int layer) { return raw_of(c, kHead + kTowers[t] + "." + std::to_string(3 * layer + 1) + ".bias"); };
    Act cur[3][5];
    for (int i = 0; i < 4; ++i) {  // cls_tower and center_tower, layer i, all levels
        Act raw[2][5];
        const bool in_raw = gn_fuse && i > 0;  // cur[t][l] is the raw output of layer i - 1
        B.begin_group(std::string("cls_tower+center_tower.") + std::to_string(3 * i));
        for (int t = 0; t < 2; ++t)
            for (int l = 0; l < 5; ++l)
                raw[t][l] = B.conv(B.layer(kHead + kTowers[t] + "." + std::to_string(3 * i)), i == 0 ? P[l] : cur[t][l],
                                   false, nullptr, 0, sums_of(t, i, l), nullptr, 0,
                                   in_raw ? sums_of(t, i - 1, l) : nullptr, in_raw && base ? gn_w(t, i - 1) : nullptr,
                                   in_raw && base ? gn_b(t, i - 1) : nullptr);
        B.end_group();
        if (B.failed) return -1;
        std::vector<std::pair<int, Act*>> items;
        for (int t = 0; t < 2; ++t)
            for (int l = 0; l < 5; ++l) {
                if (i == 0) {
                    if (t == 1) B.free_act(P[l]);
                } else {
                    B.free_act(cur[t][l]);
                }
                cur[t][l] = raw[t][l];
                items.push_back({t, &cur[t][l]});
            }
        if (!gn_fuse || i == 3) gn_group(items, i);
    }
    for (int i = 0; i < 4; ++i) {  // corners_tower on top of the center tower
        Act raw[5];
        const bool in_raw = gn_fuse && i > 0;
        B.begin_group(std::string("corners_tower.") + std::to_string(3 * i));
        for (int l = 0; l < 5; ++l)
            raw[l] = B.conv(B.layer(kHead + "corners_tower." + std::to_string(3 * i)), i == 0 ? cur[1][l] : cur[2][l],
                            false, nullptr, 0, sums_of(2, i, l), nullptr, 0,
                            in_raw ? sums_of(2, i - 1, l) : nullptr, in_raw && base ? gn_w(2, i - 1) : nullptr,
                            in_raw && base ? gn_b(2, i - 1) : nullptr);
        B.end_group();
        if (B.failed) return -1;
        std::vector<std::pair<int, Act*>> items;
        for (int l = 0; l < 5; ++l) {
            if (i > 0) B.free_act(cur[2][l]);
            cur[2][l] = raw[l];
            items.push_back({2, &cur[2][l]});
        }
        if (!gn_fuse || i == 3) gn_group(items, i);
    }
    for (int t = 0; t < 3; ++t)
        for (int l = 0; l < 5; ++l) B.name(std::string(kTowers[t]) + ".l" + std::to_string(l), cur[t][l]);
    // all levels of one prediction conv next to each other: its weights stay resident in shared memory across them
    B.begin_group("pred.cls_logits+center_pred");
    for (int l = 0; l < 5; ++l)
        B.conv(B.layer(kHead + "cls_logits"), cur[0][l], false, nullptr, 0, nullptr,
               base ? ho[l][0].p : reinterpret_cast<float*>(1), ho[l][0].ld);
    for (int l = 0; l < 5; ++l)
        B.conv(B.layer(kHead + "center_pred"), cur[1][l], false, nullptr, 0, nullptr,
               base ? ho[l][2].p : reinterpret_cast<float*>(1), ho[l][2].ld);
    B.end_group();
    B.begin_group("pred.ctrness+corners_pred");
    for (int l = 0; l < 5; ++l)
        B.conv(B.layer(kHead + "ctrness"), cur[2][l], false, nullptr, 0, nullptr,
               base ? ho[l][1].p : reinterpret_cast<float*>(1), ho[l][1].ld);
    B.end_group();
    for (int t = 0; t < 3; ++t)
        for (int l = 0; l < 5; ++l) B.free_act(cur[t][l]);
    if (B.failed) return -1;

    if (base) {
        // the kernels read their problem descriptors (tensor maps + parameters) from device memory
        CUDA_OK(cudaMemcpy(B.dev_probs, B.probs.data(), B.probs.size() * sizeof(ConvProblem), cudaMemcpyHostToDevice));
        if (!B.pairs.empty())
            CUDA_OK(cudaMemcpy(B.dev_pairs, B.pairs.data(), B.pairs.size() * sizeof(PairProblem), cudaMemcpyHostToDevice));
        if (!B.tails.empty())
            CUDA_OK(cudaMemcpy(B.dev_tails, B.tails.data(), B.tails.size() * sizeof(TailProblem), cudaMemcpyHostToDevice));
    }

    if (needed) *needed = B.arena.peak;
    if (!base) return 0;
    if (B.arena.peak > bytes) {
        set_error("bind_workspace: workspace of %zu bytes is too small, need %zu", bytes, B.arena.peak);
        c->ops.clear();
        return -1;
    }
    c->N = N;
    c->H = H;
    c->W = W;
    c->ws = base;
    c->ws_bytes = bytes;
    c->launches_per_forward = B.launches;  // preprocess + stem (with pool) + convs + GN applies + relu copy
    c->x0 = x0p_saved;
    c->flops_per_forward = B.flops;
    c->gn_sums_all = sums_all;
    c->gn_sums_bytes = sums_bytes;
    c->sizes_dev = sizes_dev;
    c->images_dev = images_dev;
    c->images_dev_bytes = img_bytes;
    c->dets_dev = dets_dev;
    c->counts_dev = counts_dev;
    c->dets_capacity = det_cap;
    c->images_dev2 = images_dev2;
    c->dets_dev2 = dets_dev2;
    c->counts_dev2 = counts_dev2;
    c->slot_pending[0] = c->slot_pending[1] = false;
    c->slot_used[0] = c->slot_used[1] = false;
    c->post_scratch = post_scratch;
    c->post_scratch_bytes = post_bytes;
    for (int l = 0; l < 5; ++l)
        for (int k = 0; k < 3; ++k) c->head_out[l][k] = ho[l][k];
    return 0;
}

// Host int32 values reach the device as kernel arguments (copied at launch, no pinned staging, no host sync).
struct IntChunk {
    int32_t v[256];
};
__global__ void set_ints_kernel(IntChunk a, int32_t* dst, int n) {
    const int i = threadIdx.x;
    if (i < n) dst[i] = a.v[i];
}

static int upload_sizes(dafne_ctx* c, const int32_t* image_sizes, const int32_t* output_sizes, cudaStream_t s) {
    for (int n0 = 0; n0 < c->N; n0 += 64) {
        IntChunk ch;
        const int cnt = c->N - n0 < 64 ? c->N - n0 : 64;
        for (int j = 0; j < cnt; ++j) {
            const int n = n0 + j;
            const int h = image_sizes ? image_sizes[2 * n] : c->H, w = image_sizes ? image_sizes[2 * n + 1] : c->W;
            if (h < 1 || w < 1 || h > c->H || w > c->W) {
                set_error("image %d size %dx%d outside the bound %dx%d batch", n, h, w, c->H, c->W);
                return -1;
            }
            ch.v[4 * j] = h;
            ch.v[4 * j + 1] = w;
            ch.v[4 * j + 2] = output_sizes ? output_sizes[2 * n] : h;
            ch.v[4 * j + 3] = output_sizes ? output_sizes[2 * n + 1] : w;
        }
        set_ints_kernel<<<1, 256, 0, s>>>(ch, c->sizes_dev + 4 * n0, 4 * cnt);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            set_error("set_ints_kernel launch: %s", cudaGetErrorString(e));
            return -1;
        }
        c->stat_launches += 1;
    }
    return 0;
}

int ctx_forward(dafne_ctx* c, const void* images, int dtype, const int32_t* image_sizes, cudaStream_t s) {
    if (!c->ws || c->ops.empty()) {
        set_error("dafne_forward_dense: no workspace bound");
        return -1;
    }
    CUDA_OK(cudaSetDevice(c->device));
    if (upload_sizes(c, image_sizes, nullptr, s)) return -1;
    CUDA_OK(cudaMemsetAsync(c->gn_sums_all, 0, c->gn_sums_bytes, s));
    const bool prof = c->profiling;
    if (prof) {
        while (c->prof_events.size() < c->ops.size() + 2) {
            cudaEvent_t e;
            CUDA_OK(cudaEventCreate(&e));
            c->prof_events.push_back(e);
        }
        CUDA_OK(cudaEventRecord(c->prof_events[0], s));
    }
    if (launch_preprocess(images, dtype, c->sizes_dev, c->N, c->H, c->W, c->spec.pixel_mean, c->spec.pixel_std,
                          c->x0, s))
        return -1;
    if (prof) CUDA_OK(cudaEventRecord(c->prof_events[1], s));
    for (size_t i = 0; i < c->ops.size(); ++i) {
        if (c->ops[i](s)) return -1;
        if (prof) CUDA_OK(cudaEventRecord(c->prof_events[i + 2], s));
    }
    c->stat_launches += c->launches_per_forward;
    c->stat_flops += c->flops_per_forward;
    return 0;
}

void fill_post_spec(const dafne_ctx* c, PostParams* p) {
    const dafne_model_spec& sp = c->spec;
    p->L = sp.num_levels;
    p->num_classes = sp.num_classes;
    p->sort_corners = sp.sort_corners;
    p->thresh_with_ctr = sp.thresh_with_ctr;
    p->pre_nms_topk = sp.pre_nms_topk;
    p->post_nms_topk = sp.post_nms_topk;
    p->vehicle_merge = sp.vehicle_merge;
    p->score_thresh = sp.score_thresh;
    p->nms_thresh = sp.nms_thresh;
    for (int l = 0; l < sp.num_levels; ++l) p->lv[l].stride = sp.fpn_strides[l];
}

int ctx_postprocess(dafne_ctx* c, const int32_t* image_sizes, const int32_t* output_sizes, int do_postprocess,
                    float* dets, int32_t* counts, int capacity, cudaStream_t s) {
    if (!c->ws) {
        set_error("dafne_postprocess: no workspace bound");
        return -1;
    }
    CUDA_OK(cudaSetDevice(c->device));
    if (upload_sizes(c, image_sizes, output_sizes, s)) return -1;
    PostParams p;
    memset(&p, 0, sizeof(p));
    fill_post_spec(c, &p);
    p.N = c->N;
    for (int l = 0; l < 5; ++l) {
        PostLevel& lv = p.lv[l];
        lv.logits = c->head_out[l][0].p;
        lv.ld_logits = c->head_out[l][0].ld;
        lv.ctr = c->head_out[l][1].p;  // column 0 = ctrness logit
        lv.ld_ctr = c->head_out[l][1].ld;
        lv.reg = c->head_out[l][1].p + 1;  // columns 1..8 = corner deltas
        lv.ld_reg = c->head_out[l][1].ld;
        lv.center = c->head_out[l][2].p;
        lv.ld_center = c->head_out[l][2].ld;
        lv.H = c->head_out[l][0].H;
        lv.W = c->head_out[l][0].W;
    }
    p.scales_dev = c->scales_dev;
    p.do_postprocess = do_postprocess;
    p.sizes_dev = c->sizes_dev;
    p.dets = dets;
    p.counts = counts;
    p.capacity = capacity;
    p.scratch = c->post_scratch;
    p.scratch_bytes = c->post_scratch_bytes;
    return launch_postprocess(p, s, &c->stat_launches);
}

}  // namespace dafne
