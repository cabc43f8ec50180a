// Rotated-box post-processing of DAFNe on device, with no host round trip:
//   score = sqrt(sigmoid(cls) * sigmoid(ctr)), threshold, per-level top-k     dafne/modeling/dafne/dafne_outputs.py:792-858
//   center-to-corner decode, polygon = location + reg * stride                 dafne/modeling/dafne/dafne.py:405-411, dafne_outputs.py:771-772,860-872
//   canonical corner order                                                     dafne/utils/sort_corners.py:26-92
//   class-aware polygon NMS (class offset trick, 5->4 merge, IoU > thr)        dafne/modeling/nms/nms.py:37-92 -> external poly_gpu_nms
//   polygon IoU arithmetic (fp32 transliteration of the in-tree algorithm)     tools/prepare_dota/polyiou.cpp:8-133
//   post-NMS top-k with ties kept                                              dafne/modeling/dafne/dafne_outputs.py:907-925
//   rescale / clip / drop empty                                                detectron2 detector_postprocess, one_stage_detector.py:78-98
//
// This translation unit is compiled with -fmad=false -prec-div=true -prec-sqrt=true -ftz=false: every fp32 operation
// is rounded once, in source order, exactly like the CPU oracle (oracle/polyiou_oracle.c, oracle/postprocess.py).
// Sigmoid is evaluated in double and rounded once to float (the correctly rounded value; torch's float sigmoid is
// within 1-2 ulp of it and differs between its own CPU and CUDA builds).
//
// Ordering contract (the reference leaves ties unspecified): candidates are ordered by descending score, ties by
// ascending canonical index (level, location, class).
#include "postprocess.cuh"

#include <math.h>
#include <stdio.h>

#include "conv_tc.cuh"  // set_error
#include "polyiou.cuh"

namespace dafne {

#define POST_CHECK_LAUNCH(name)                                         \
    do {                                                                \
        cudaError_t e__ = cudaGetLastError();                           \
        if (e__ != cudaSuccess) {                                       \
            set_error("%s launch: %s", name, cudaGetErrorString(e__)); \
            return -1;                                                  \
        }                                                               \
    } while (0)

constexpr int kMaxSorted = 16384;  // candidates per image the in-smem sort handles
constexpr int kDet = 20;

// polygon IoU arithmetic: polyiou.cuh (shared with nms.cu)
__global__ void poly_iou_kernel(const float* __restrict__ p, const float* __restrict__ q, float* __restrict__ out,
                                int n) {
    __shared__ float2 slots[9 * 64];  // launched with 64 threads: one private column of 9 points per thread
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = iou_poly_f32(p + 8 * i, q + 8 * i, slots + threadIdx.x, 64);
}
__global__ void pair_filter_kernel(const float* __restrict__ p, const float* __restrict__ q,
                                   unsigned char* __restrict__ fired, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float a[8], b[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        a[k] = p[8 * i + k];
        b[k] = q[8 * i + k];
    }
    const NmsAux P = nms_aux_of(a), Q = nms_aux_of(b);
    P2 pp[6], qq[6];
    load_oriented(a, pp);
    load_oriented(b, qq);
    float po[8], qo[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        po[2 * k] = pp[k].x;
        po[2 * k + 1] = pp[k].y;
        qo[2 * k] = qq[k].x;
        qo[2 * k + 1] = qq[k].y;
    }
    fired[i] = (pair_inter_is_zero(P, Q, po, qo) && (P.area + Q.area) != 0.f) ? 1 : 0;
}
// Test hook for the per-term filter (polyiou.cuh::term_is_zero): per pair, bit k = 4 * i + j of `fired` says the filter
// declared term (edge i of P, edge j of Q) zero, the same bit of `nonzero` that tri_overlap's value is not (+-)0.
__global__ void term_filter_kernel(const float* __restrict__ p, const float* __restrict__ q,
                                   unsigned short* __restrict__ fired, unsigned short* __restrict__ nonzero, int n) {
    __shared__ float2 slots[9 * 64];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    float a[8], b[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        a[k] = p[8 * idx + k];
        b[k] = q[8 * idx + k];
    }
    const NmsAux P = nms_aux_of(a), Q = nms_aux_of(b);
    const TermPairCtx ctx = term_pair_ctx(P, Q);
    P2 pp[6], qq[6];
    load_oriented(a, pp);
    load_oriented(b, qq);
    unsigned f = 0, nz = 0;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            const P2 A = pp[i], B = pp[i + 1], Cc = qq[j], D = qq[j + 1];
            if (term_is_zero(ctx, A, B, Cc, D, vertex_t(A.x, A.y), vertex_t(B.x, B.y), vertex_t(Cc.x, Cc.y),
                             vertex_t(D.x, D.y)))
                f |= 1u << (4 * i + j);
            if (tri_overlap(A, B, Cc, D, slots + threadIdx.x, 64) != 0.f) nz |= 1u << (4 * i + j);
        }
    fired[idx] = static_cast<unsigned short>(f);
    nonzero[idx] = static_cast<unsigned short>(nz);
}
int launch_term_filter(const float* p, const float* q, unsigned short* fired, unsigned short* nonzero, int n,
                       cudaStream_t s) {
    if (n <= 0) return 0;
    term_filter_kernel<<<(n + 63) / 64, 64, 0, s>>>(p, q, fired, nonzero, n);
    POST_CHECK_LAUNCH("term_filter_kernel");
    return 0;
}
int launch_pair_filter(const float* p, const float* q, unsigned char* fired, int n, cudaStream_t s) {
    if (n <= 0) return 0;
    pair_filter_kernel<<<(n + 127) / 128, 128, 0, s>>>(p, q, fired, n);
    POST_CHECK_LAUNCH("pair_filter_kernel");
    return 0;
}
int launch_poly_iou(const float* p, const float* q, float* iou, int n, cudaStream_t s) {
    if (n <= 0) return 0;
    poly_iou_kernel<<<(n + 63) / 64, 64, 0, s>>>(p, q, iou, n);
    POST_CHECK_LAUNCH("poly_iou_kernel");
    return 0;
}

// ================================================================================================ corner sort
__device__ __forceinline__ float cross2(float ax, float ay, float bx, float by) { return ax * by - ay * bx; }

// sort_corners.py:26-92, one quadrilateral per thread. Pure permutation of the inputs, or zeros for degenerate rows.
__device__ __forceinline__ void sort_quad(const float* in, float* out) {
    float px[4], py[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        px[i] = in[2 * i];
        py[i] = in[2 * i + 1];
    }
    int lm = 0;  // leftmost vertex, first minimum wins (:46)
#pragma unroll
    for (int i = 1; i < 4; ++i)
        if (px[i] < px[lm]) lm = i;
    const float p1x = px[lm], p1y = py[lm];
    float sx[3], sy[3];
    {
        int k = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i != lm) {
                sx[k] = px[i];
                sy[k] = py[i];
                ++k;
            }
    }
    float p3x = 0.f, p3y = 0.f, ax = 0.f, ay = 0.f, bx = 0.f, by = 0.f;  // (a, b) = S_new
    bool done = false;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int j = i == 0 ? 1 : 0;
        const int k = i == 2 ? 1 : 2;
        const float l = cross2(sx[i] - p1x, sy[i] - p1y, sx[j] - p1x, sy[j] - p1y);
        const float r = cross2(sx[i] - p1x, sy[i] - p1y, sx[k] - p1x, sy[k] - p1y);
        if (!done && (l * r) < 0.0f) {
            p3x = sx[i];
            p3y = sy[i];
            ax = sx[j];
            ay = sy[j];
            bx = sx[k];
            by = sy[k];
            done = true;
        }
    }
    float p2x, p2y, p4x, p4y;
    const float c0 = cross2(p3x - p1x, p3y - p1y, ax - p1x, ay - p1y);
    const float c1 = cross2(p3x - p1x, p3y - p1y, bx - p1x, by - p1y);
    if (c0 > 0.0f || !(c1 > 0.0f)) {
        p2x = ax;
        p2y = ay;
        p4x = bx;
        p4y = by;
    } else {
        p2x = bx;
        p2y = by;
        p4x = ax;
        p4y = ay;
    }
    out[0] = p1x;
    out[1] = p1y;
    out[2] = p2x;
    out[3] = p2y;
    out[4] = p3x;
    out[5] = p3y;
    out[6] = p4x;
    out[7] = p4y;
}

__global__ void sort_quad_kernel(const float* __restrict__ in, float* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float q[8], o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) q[k] = in[8 * i + k];
    sort_quad(q, o);
#pragma unroll
    for (int k = 0; k < 8; ++k) out[8 * i + k] = o[k];
}
int launch_sort_quadrilateral(const float* quads, float* out, int n, cudaStream_t s) {
    if (n <= 0) return 0;
    sort_quad_kernel<<<(n + 127) / 128, 128, 0, s>>>(quads, out, n);
    POST_CHECK_LAUNCH("sort_quad_kernel");
    return 0;
}

// ================================================================================================ scratch layout
struct Layout {
    int N, L, C, ldc;
    int hw[5], cap[5];
    size_t cand_off[5];  // offsets (in keys) of each level's candidate list inside one image's block
    size_t cand_per_img;  // keys per image
    int level_base[6];    // canonical index base of each level (sum of HW*C of the lower levels)
    int max_sel, nblk;
    // byte offsets of the per-image arrays, each [N][...]
    size_t o_cand, o_cand_cnt, o_sel, o_sel_cnt, o_poly, o_nmsbox, o_score, o_ctr, o_cls, o_level, o_loc, o_hbox,
        o_canon, o_nms, o_keep, o_nkeep, total;
    size_t nms_bytes;
};

static size_t a256(size_t v) { return (v + 255) / 256 * 256; }

static Layout make_layout(int N, int L, const int* level_hw, int C, int topk) {
    Layout y;
    y.N = N;
    y.L = L;
    y.C = C;
    y.ldc = C <= 16 ? 16 : 32;
    size_t keys = 0;
    int sel = 0, base = 0;
    for (int l = 0; l < L; ++l) {
        y.hw[l] = level_hw[2 * l] * level_hw[2 * l + 1];
        y.cap[l] = y.hw[l] * C;
        y.cand_off[l] = keys;
        keys += y.cap[l];
        y.level_base[l] = base;
        base += y.cap[l];
        sel += y.cap[l] < topk ? y.cap[l] : topk;
    }
    y.level_base[L] = base;
    y.cand_per_img = keys;
    y.max_sel = sel < 1 ? 1 : sel;
    y.nblk = (y.max_sel + 63) / 64;
    size_t o = 0;
    const size_t n = N, ms = y.max_sel;
    y.o_cand = o;
    o = a256(o + n * keys * 8);
    y.o_cand_cnt = o;
    o = a256(o + n * 8 * 4);
    y.o_sel = o;
    o = a256(o + n * ms * 8);
    y.o_sel_cnt = o;
    o = a256(o + n * 4);
    y.o_poly = o;
    o = a256(o + n * ms * 32);
    y.o_nmsbox = o;
    o = a256(o + n * ms * 32);
    y.o_score = o;
    o = a256(o + n * ms * 4);
    y.o_ctr = o;
    o = a256(o + n * ms * 4);
    y.o_cls = o;
    o = a256(o + n * ms * 4);
    y.o_level = o;
    o = a256(o + n * ms * 4);
    y.o_loc = o;
    o = a256(o + n * ms * 8);
    y.o_hbox = o;
    o = a256(o + n * ms * 16);
    y.o_canon = o;
    o = a256(o + n * ms * 4);
    y.o_nms = o;
    y.nms_bytes = nms_scratch_bytes(N, y.max_sel);
    o = a256(o + y.nms_bytes);
    y.o_keep = o;
    o = a256(o + n * ms * 4);
    y.o_nkeep = o;
    o = a256(o + n * 4);
    y.total = o;
    return y;
}

size_t postprocess_scratch_bytes(int N, int L, const int* level_hw, int num_classes, int pre_nms_topk) {
    return make_layout(N, L, level_hw, num_classes, pre_nms_topk).total;
}

// ================================================================================================ K1: scores + candidates
__device__ __forceinline__ float sigmoid_cr(float x) { return static_cast<float>(1.0 / (1.0 + exp(-static_cast<double>(x)))); }

struct ScoreArgs {
    const float* logits;
    const float* ctr;
    int ldc, ld_logits, ld_ctr, C, HW;
    float thr;
    int thresh_with_ctr;
    unsigned long long* cand;  // [N][cand_per_img] + level offset
    size_t cand_per_img;
    int* cand_cnt;  // [N][8] + level
};

// one thread per (image, location, padded class); ldc in {16, 32} so a location never straddles a warp
__global__ void score_candidates_kernel(ScoreArgs a, int N) {
    const size_t total = static_cast<size_t>(N) * a.HW * a.ldc;
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    const unsigned lane = threadIdx.x & 31;
    bool cand = false;
    unsigned long long key = 0;
    int n = 0;
    if (i < total) {
        const int c = i % a.ldc;
        const size_t pix = i / a.ldc;  // n * HW + loc
        n = pix / a.HW;
        const int loc = pix % a.HW;
        if (c < a.C) {
            const float cls = sigmoid_cr(a.logits[pix * a.ld_logits + c]);
            const float ctr = sigmoid_cr(a.ctr[pix * a.ld_ctr]);
            float s;
            if (a.thresh_with_ctr) {
                s = sqrtf(cls * ctr);
                cand = s > a.thr;
            } else {
                cand = cls > a.thr;
                s = sqrtf(cls * ctr);
            }
            const unsigned idx = static_cast<unsigned>(loc) * a.C + c;
            key = (static_cast<unsigned long long>(__float_as_uint(s)) << 32) | (0xFFFFFFFFu - idx);
        }
    }
    // warp-aggregated append; a warp can straddle two images only at an image boundary, handle both
    const int n_first = __shfl_sync(0xffffffffu, n, 0);
    for (int pass = 0; pass < 2; ++pass) {
        const int img = n_first + pass;
        const unsigned m = __ballot_sync(0xffffffffu, cand && n == img);
        if (m == 0) continue;
        const int leader = __ffs(m) - 1;
        int basepos = 0;
        if (lane == leader) basepos = atomicAdd(a.cand_cnt + img * 8, __popc(m));
        basepos = __shfl_sync(0xffffffffu, basepos, leader);
        if (cand && n == img) {
            const int pos = basepos + __popc(m & ((1u << lane) - 1));
            a.cand[static_cast<size_t>(img) * a.cand_per_img + pos] = key;
        }
    }
}

// ================================================================================================ K2: per-level top-k
struct SelectArgs {
    const unsigned long long* cand;  // image 0, level 0
    size_t cand_per_img;
    size_t cand_off[5];
    const int* cand_cnt;  // [N][8]
    int level_base[5];
    int topk;
    unsigned long long* sel;  // [N][max_sel]
    int* sel_cnt;             // [N]
    int max_sel;
};

// one CTA per (level, image): exact radix select of the topk largest 64-bit keys (all keys are distinct)
__global__ void __launch_bounds__(1024) select_topk_kernel(SelectArgs a) {
    const int l = blockIdx.x, n = blockIdx.y;
    const int cnt = a.cand_cnt[n * 8 + l];
    const unsigned long long* keys = a.cand + static_cast<size_t>(n) * a.cand_per_img + a.cand_off[l];
    __shared__ unsigned hist[256];
    __shared__ unsigned long long s_prefix;
    __shared__ int s_k, s_base;
    unsigned long long thr_key = 0;  // keep keys >= thr_key
    if (cnt > a.topk) {
        if (threadIdx.x == 0) {
            s_prefix = 0;
            s_k = a.topk;
        }
        for (int byte = 7; byte >= 0; --byte) {
            for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
            __syncthreads();
            const unsigned long long prefix = s_prefix;
            const int shift = 8 * (byte + 1);
            for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
                const unsigned long long k = keys[i];
                const bool match = byte == 7 ? true : ((k >> shift) == prefix);
                if (match) atomicAdd(&hist[(k >> (8 * byte)) & 0xFF], 1u);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                int k = s_k;
                int b = 255;
                for (; b > 0; --b) {
                    if (static_cast<int>(hist[b]) >= k) break;
                    k -= hist[b];
                }
                s_k = k;
                s_prefix = (prefix << 8) | static_cast<unsigned>(b);
            }
            __syncthreads();
        }
        thr_key = s_prefix;
    }
    const int take = cnt > a.topk ? a.topk : cnt;
    if (threadIdx.x == 0) s_base = atomicAdd(a.sel_cnt + n, take);
    __syncthreads();
    // append (order inside the image list is irrelevant: the list is sorted next)
    __shared__ int s_pos;
    if (threadIdx.x == 0) s_pos = 0;
    __syncthreads();
    unsigned long long* dst = a.sel + static_cast<size_t>(n) * a.max_sel + s_base;
    const unsigned lb = a.level_base[l];
    for (int i0 = 0; i0 < cnt; i0 += blockDim.x) {
        const int i = i0 + threadIdx.x;
        unsigned long long k = 0;
        bool ok = false;
        if (i < cnt) {
            k = keys[i];
            ok = k >= thr_key;
        }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        const unsigned lane = threadIdx.x & 31;
        int basepos = 0;
        if (m != 0) {
            const int leader = __ffs(m) - 1;
            if (lane == leader) basepos = atomicAdd(&s_pos, __popc(m));
            basepos = __shfl_sync(0xffffffffu, basepos, leader);
        }
        if (ok) {
            const unsigned idx = 0xFFFFFFFFu - static_cast<unsigned>(k & 0xFFFFFFFFu);  // index inside the level
            const unsigned canon = lb + idx;
            dst[basepos + __popc(m & ((1u << lane) - 1))] = (k & 0xFFFFFFFF00000000ull) | (0xFFFFFFFFu - canon);
        }
    }
}

// ================================================================================================ K3: sort + decode
struct DecodeArgs {
    PostLevel lv[5];
    int L, C;
    int level_base[6];
    const float* scales;
    int sort_corners, vehicle_merge;
    const unsigned long long* sel;
    const int* sel_cnt;
    int max_sel;
    float *poly, *nmsbox, *score, *ctr, *loc, *hbox;
    int *cls, *level;
    unsigned* canon;
    unsigned* minmax;  // [N][8] view of cand_cnt: slots 6 / 7 = ~key(min), key(max) of the image's coordinates
};

__device__ __forceinline__ void bitonic_sort_desc(unsigned long long* s, int n_pow2) {
    for (int k = 2; k <= n_pow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = s[i], b = s[ixj];
                    const bool desc = (i & k) == 0;
                    if (desc ? (a < b) : (a > b)) {
                        s[i] = b;
                        s[ixj] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
}

// Monotonic float -> unsigned map (order of the reals, -0 < +0), so block results can be merged with integer atomicMax.
__device__ __forceinline__ unsigned float_order_key(float v) {
    const unsigned b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float float_from_order_key(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

// Sort by rank + decode. The selected keys of an image are distinct (the canonical index is in the low word), so the
// position of a key in the descending order is the number of larger keys: every thread counts that for ONE key against
// the image's key list in shared memory (all lanes read the same address: a broadcast) and then decodes its candidate
// straight into row `rank`. A block takes kRankRows keys, so an image is spread over ceil(m / kRankRows) CTAs and a
// batch over all SMs; the one-CTA-per-image bitonic sort this replaces kept 8 SMs busy for 90 us. The min / max of all
// coordinates of the image (nms.py:74-75) is merged with two integer atomicMax on order-preserving keys.
constexpr int kRankRows = 256;

__global__ void __launch_bounds__(kRankRows) rank_decode_kernel(DecodeArgs a) {
    extern __shared__ __align__(16) unsigned long long s_keys[];
    __shared__ float s_min[kRankRows / 32], s_max[kRankRows / 32];
    const int n = blockIdx.y;
    const int m = a.sel_cnt[n];
    const int i0 = blockIdx.x * kRankRows;
    if (i0 >= m) return;
    const unsigned long long* src = a.sel + static_cast<size_t>(n) * a.max_sel;
    for (int i = threadIdx.x; i < m; i += kRankRows) s_keys[i] = src[i];
    __syncthreads();
    const int i = i0 + threadIdx.x;
    const bool on = i < m;
    const unsigned long long key = on ? s_keys[i] : 0ull;
    int r = 0;
    {
        // two keys per 16-byte broadcast load (the dynamic shared-memory window is 16-byte aligned)
        const ulonglong2* k2 = reinterpret_cast<const ulonglong2*>(s_keys);
        int j = 0;
#pragma unroll 4
        for (; j + 4 <= m; j += 4) {
            const ulonglong2 u = k2[j >> 1], v = k2[(j >> 1) + 1];
            r += (u.x > key) + (u.y > key) + (v.x > key) + (v.y > key);
        }
        for (; j < m; ++j) r += s_keys[j] > key;
    }

    const size_t rowbase = static_cast<size_t>(n) * a.max_sel;
    float vmin = INFINITY, vmax = -INFINITY;
    if (on) {
        const float score = __uint_as_float(static_cast<unsigned>(key >> 32));
        const unsigned canon = 0xFFFFFFFFu - static_cast<unsigned>(key & 0xFFFFFFFFu);
        int l = 0;
        while (l + 1 < a.L && canon >= static_cast<unsigned>(a.level_base[l + 1])) ++l;
        const unsigned idx = canon - a.level_base[l];
        const int loc = idx / a.C, c = idx % a.C;
        const PostLevel& lv = a.lv[l];
        const size_t pix = static_cast<size_t>(n) * lv.H * lv.W + loc;
        const int gx = loc % lv.W, gy = loc / lv.W;
        const float fs = static_cast<float>(lv.stride);
        const float lx = static_cast<float>(gx * lv.stride) + static_cast<float>(lv.stride / 2);
        const float ly = static_cast<float>(gy * lv.stride) + static_cast<float>(lv.stride / 2);
        float q[8], o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float reg = lv.reg[pix * lv.ld_reg + k];
            if (lv.center != nullptr) {
                // reg_corners = (center.repeat(1,4,1,1) + delta) * scale_l      (dafne.py:405-411)
                reg = (lv.center[pix * lv.ld_center + (k & 1)] + reg) * a.scales[l];
            }
            const float rc = reg * fs;       // dafne_outputs.py:771-772
            q[k] = ((k & 1) ? ly : lx) + rc;  // dafne_outputs.py:860-872
        }
        if (a.sort_corners) {
            sort_quad(q, o);
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = q[k];
        }
        float xmin = o[0], xmax = o[0], ymin = o[1], ymax = o[1];
#pragma unroll
        for (int k = 1; k < 4; ++k) {
            xmin = fminf(xmin, o[2 * k]);
            xmax = fmaxf(xmax, o[2 * k]);
            ymin = fminf(ymin, o[2 * k + 1]);
            ymax = fmaxf(ymax, o[2 * k + 1]);
        }
        vmin = fminf(xmin, ymin);
        vmax = fmaxf(xmax, ymax);
        float4* pp = reinterpret_cast<float4*>(a.poly + (rowbase + r) * 8);
        pp[0] = make_float4(o[0], o[1], o[2], o[3]);
        pp[1] = make_float4(o[4], o[5], o[6], o[7]);
        *reinterpret_cast<float4*>(a.hbox + (rowbase + r) * 4) = make_float4(xmin, ymin, xmax, ymax);
        a.score[rowbase + r] = score;
        a.ctr[rowbase + r] = sigmoid_cr(lv.ctr[pix * lv.ld_ctr]);
        a.cls[rowbase + r] = c;
        a.level[rowbase + r] = l;
        *reinterpret_cast<float2*>(a.loc + (rowbase + r) * 2) = make_float2(lx, ly);
        a.canon[rowbase + r] = canon;
    }
    // min / max of all coordinates of the image (nms.py:74-75): warp, block, then one atomic pair per block
    for (int o = 16; o > 0; o >>= 1) {
        vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
        vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    }
    if ((threadIdx.x & 31) == 0) {
        s_min[threadIdx.x >> 5] = vmin;
        s_max[threadIdx.x >> 5] = vmax;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kRankRows / 32; ++w) {
            vmin = fminf(vmin, s_min[w]);
            vmax = fmaxf(vmax, s_max[w]);
        }
        atomicMax(a.minmax + n * 8 + 6, ~float_order_key(vmin));  // both slots start at 0 (cleared with cand_cnt)
        atomicMax(a.minmax + n * 8 + 7, float_order_key(vmax));
    }
}

// boxes handed to the NMS: every class shifted by class * (max - min + 1) in fp32 (nms.py:74-83)
__global__ void __launch_bounds__(256) nms_offset_kernel(DecodeArgs a) {
    const int n = blockIdx.y;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.sel_cnt[n]) return;
    const float vmin = float_from_order_key(~a.minmax[n * 8 + 6]);
    const float vmax = float_from_order_key(a.minmax[n * 8 + 7]);
    const float span = vmax - vmin + 1.0f;  // max_coordinate - min_coordinate + 1   (nms.py:81)
    const size_t row = static_cast<size_t>(n) * a.max_sel + r;
    int c = a.cls[row];
    if (a.vehicle_merge && c == 5) c = 4;  // nms.py:77-79
    const float off = static_cast<float>(c) * span;
    const float4* pp = reinterpret_cast<const float4*>(a.poly + row * 8);
    float4* nb = reinterpret_cast<float4*>(a.nmsbox + row * 8);
    const float4 p0 = pp[0], p1 = pp[1];
    nb[0] = make_float4(p0.x + off, p0.y + off, p0.z + off, p0.w + off);
    nb[1] = make_float4(p1.x + off, p1.y + off, p1.z + off, p1.w + off);
}

// K4: lazily evaluated NMS -> nms.cu (run_nms)

// ================================================================================================ K5: top-k cut, rescale, clip, pack
struct FinalArgs {
    const float *poly, *score, *ctr, *loc, *hbox;
    const int *cls, *level;
    const unsigned* canon;
    const int *keep, *nkeep;
    int max_sel, post_topk, do_postprocess;
    const int32_t* sizes;  // [N][4] h, w, out_h, out_w
    float* dets;
    int32_t* counts;
    int capacity;
};

__global__ void __launch_bounds__(1024) finalize_kernel(FinalArgs a) {
    const int n = blockIdx.x;
    const size_t rowbase = static_cast<size_t>(n) * a.max_sel;
    const int* kp = a.keep + rowbase;
    int nk = a.nkeep[n];
    __shared__ int s_cut, s_out;
    if (threadIdx.x == 0) {
        int cut = nk;
        if (a.post_topk > 0 && nk > a.post_topk) {
            // kthvalue(scores, nk - topk + 1) = the topk-th largest; keep scores >= it (ties kept)  (dafne_outputs.py:916-923)
            const float thr = a.score[rowbase + kp[a.post_topk - 1]];
            cut = a.post_topk;
            while (cut < nk && a.score[rowbase + kp[cut]] >= thr) ++cut;
        }
        s_cut = cut;
        s_out = 0;
    }
    __syncthreads();
    nk = s_cut;
    const float ih = static_cast<float>(a.sizes[4 * n]), iw = static_cast<float>(a.sizes[4 * n + 1]);
    const int oh_i = a.sizes[4 * n + 2], ow_i = a.sizes[4 * n + 3];
    const float sx = static_cast<float>(static_cast<double>(ow_i) / static_cast<double>(iw));
    const float sy = static_cast<float>(static_cast<double>(oh_i) / static_cast<double>(ih));
    const float ow = static_cast<float>(ow_i), oh = static_cast<float>(oh_i);
    float* out = a.dets + static_cast<size_t>(n) * a.capacity * kDet;
    // order-preserving compaction, one block-wide round per blockDim.x rows (the usual <= ~1000 kept rows: one round)
    for (int r0 = 0; r0 < nk; r0 += blockDim.x) {
        const int r = r0 + threadIdx.x;
        bool ok = false;
        float hb[4] = {0, 0, 0, 0};
        int src = 0;
        if (r < nk) {
            src = kp[r];
            const float* h = a.hbox + (rowbase + src) * 4;
            hb[0] = h[0];
            hb[1] = h[1];
            hb[2] = h[2];
            hb[3] = h[3];
            // Boxes.scale, Boxes.clip(output size), Boxes.nonempty(): detectron2's ProposalNetwork.forward runs
            // detector_postprocess on every result, whatever OneStageDetector.forward's do_postprocess says -- that
            // flag only gates the corner / location rescale below (one_stage_detector.py:45-55, 78-98)
            hb[0] = fminf(fmaxf(hb[0] * sx, 0.f), ow);
            hb[2] = fminf(fmaxf(hb[2] * sx, 0.f), ow);
            hb[1] = fminf(fmaxf(hb[1] * sy, 0.f), oh);
            hb[3] = fminf(fmaxf(hb[3] * sy, 0.f), oh);
            ok = (hb[2] - hb[0]) > 0.f && (hb[3] - hb[1]) > 0.f;
        }
        // block-wide exclusive scan of ok
        __shared__ int s_warp[32];
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        if (lane == 0) s_warp[w] = __popc(m);
        __syncthreads();
        int before = s_out;
        for (unsigned i = 0; i < w; ++i) before += s_warp[i];
        const int pos = before + __popc(m & ((1u << lane) - 1));
        if (ok && pos < a.capacity) {
            float* d = out + static_cast<size_t>(pos) * kDet;
            const float* pp = a.poly + (rowbase + src) * 8;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float v = pp[k];
                if (a.do_postprocess) v = v * ((k & 1) ? sy : sx);  // one_stage_detector.py:92-93
                d[k] = v;
            }
            d[8] = hb[0];
            d[9] = hb[1];
            d[10] = hb[2];
            d[11] = hb[3];
            d[12] = a.score[rowbase + src];
            d[13] = a.ctr[rowbase + src];
            d[14] = static_cast<float>(a.cls[rowbase + src]);
            d[15] = static_cast<float>(a.level[rowbase + src]);
            float lx = a.loc[(rowbase + src) * 2], ly = a.loc[(rowbase + src) * 2 + 1];
            if (a.do_postprocess) {
                lx = lx * sx;  // one_stage_detector.py:94-95
                ly = ly * sy;
            }
            d[16] = lx;
            d[17] = ly;
            d[18] = __uint_as_float(a.canon[rowbase + src]);  // raw bits of the canonical index
            d[19] = 0.f;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (unsigned i = 0; i < (blockDim.x >> 5); ++i) tot += s_warp[i];
            s_out += tot;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) a.counts[n] = s_out;
    // rows past the last detection are zero: the output is a fixed-shape wire record (one all-gather per batch,
    // dafne_b200/distributed.py) and the caller never has to clear it
    const int used = min(s_out, a.capacity);
    float4* z = reinterpret_cast<float4*>(out + static_cast<size_t>(used) * kDet);  // kDet * 4 B = 80 B rows: 16-B aligned
    const int nz = (a.capacity - used) * (kDet / 4);
    for (int i = threadIdx.x; i < nz; i += blockDim.x) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// ================================================================================================ host orchestration
int launch_postprocess(const PostParams& p, cudaStream_t s, int64_t* launches) {
    int level_hw[10];
    for (int l = 0; l < p.L; ++l) {
        level_hw[2 * l] = p.lv[l].H;
        level_hw[2 * l + 1] = p.lv[l].W;
    }
    const Layout y = make_layout(p.N, p.L, level_hw, p.num_classes, p.pre_nms_topk);
    if (y.total > p.scratch_bytes) {
        set_error("postprocess: scratch of %zu bytes is too small, need %zu", p.scratch_bytes, y.total);
        return -1;
    }
    if (y.max_sel > kMaxSorted) {
        set_error("postprocess: %d candidates per image exceed the supported %d", y.max_sel, kMaxSorted);
        return -1;
    }
    uint8_t* base = static_cast<uint8_t*>(p.scratch);
    auto* cand = reinterpret_cast<unsigned long long*>(base + y.o_cand);
    int* cand_cnt = reinterpret_cast<int*>(base + y.o_cand_cnt);
    auto* sel = reinterpret_cast<unsigned long long*>(base + y.o_sel);
    int* sel_cnt = reinterpret_cast<int*>(base + y.o_sel_cnt);
    // cand_cnt and sel_cnt are adjacent enough to clear separately
    cudaError_t e = cudaMemsetAsync(cand_cnt, 0, static_cast<size_t>(p.N) * 8 * 4, s);
    if (e == cudaSuccess) e = cudaMemsetAsync(sel_cnt, 0, static_cast<size_t>(p.N) * 4, s);
    if (e != cudaSuccess) {
        set_error("postprocess memset: %s", cudaGetErrorString(e));
        return -1;
    }
    for (int l = 0; l < p.L; ++l) {
        if (p.lv[l].ld_logits < p.num_classes) {
            set_error("postprocess: level %d logits pitch %d < num_classes %d", l, p.lv[l].ld_logits, p.num_classes);
            return -1;
        }
        ScoreArgs a;
        a.logits = p.lv[l].logits;
        a.ctr = p.lv[l].ctr;
        a.ldc = y.ldc;
        a.ld_logits = p.lv[l].ld_logits;
        a.ld_ctr = p.lv[l].ld_ctr;
        a.C = p.num_classes;
        a.HW = y.hw[l];
        a.thr = p.score_thresh;
        a.thresh_with_ctr = p.thresh_with_ctr;
        a.cand = cand + y.cand_off[l];
        a.cand_per_img = y.cand_per_img;
        a.cand_cnt = cand_cnt + l;
        const size_t total = static_cast<size_t>(p.N) * y.hw[l] * y.ldc;
        if (total == 0) continue;
        score_candidates_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(a, p.N);
        POST_CHECK_LAUNCH("score_candidates_kernel");
        if (launches) *launches += 1;
    }
    {
        SelectArgs a;
        a.cand = cand;
        a.cand_per_img = y.cand_per_img;
        for (int l = 0; l < 5; ++l) {
            a.cand_off[l] = l < p.L ? y.cand_off[l] : 0;
            a.level_base[l] = l < p.L ? y.level_base[l] : 0;
        }
        a.cand_cnt = cand_cnt;
        a.topk = p.pre_nms_topk;
        a.sel = sel;
        a.sel_cnt = sel_cnt;
        a.max_sel = y.max_sel;
        select_topk_kernel<<<dim3(p.L, p.N), 1024, 0, s>>>(a);
        POST_CHECK_LAUNCH("select_topk_kernel");
        if (launches) *launches += 1;
    }
    float* poly = reinterpret_cast<float*>(base + y.o_poly);
    float* nmsbox = reinterpret_cast<float*>(base + y.o_nmsbox);
    float* score = reinterpret_cast<float*>(base + y.o_score);
    float* ctr = reinterpret_cast<float*>(base + y.o_ctr);
    int* cls = reinterpret_cast<int*>(base + y.o_cls);
    int* level = reinterpret_cast<int*>(base + y.o_level);
    float* loc = reinterpret_cast<float*>(base + y.o_loc);
    float* hbox = reinterpret_cast<float*>(base + y.o_hbox);
    unsigned* canon = reinterpret_cast<unsigned*>(base + y.o_canon);
    void* nms_scratch = base + y.o_nms;
    int* keep = reinterpret_cast<int*>(base + y.o_keep);
    int* nkeep = reinterpret_cast<int*>(base + y.o_nkeep);
    {
        DecodeArgs a;
        for (int l = 0; l < 5; ++l) a.lv[l] = p.lv[l < p.L ? l : 0];
        a.L = p.L;
        a.C = p.num_classes;
        for (int l = 0; l <= p.L; ++l) a.level_base[l] = y.level_base[l];
        a.scales = p.scales_dev;
        a.sort_corners = p.sort_corners;
        a.vehicle_merge = p.vehicle_merge;
        a.sel = sel;
        a.sel_cnt = sel_cnt;
        a.max_sel = y.max_sel;
        a.poly = poly;
        a.nmsbox = nmsbox;
        a.score = score;
        a.ctr = ctr;
        a.loc = loc;
        a.hbox = hbox;
        a.cls = cls;
        a.level = level;
        a.canon = canon;
        a.minmax = reinterpret_cast<unsigned*>(cand_cnt);
        const size_t smem = static_cast<size_t>(y.max_sel) * 8;
        static DeviceOnce configured;  // largest size configured so far, per device
        int dev_rd;
        if (static_cast<int>(smem) > configured.get(&dev_rd)) {
            e = cudaFuncSetAttribute(rank_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem));
            if (e != cudaSuccess) {
                set_error("rank_decode_kernel smem attribute (%zu): %s", smem, cudaGetErrorString(e));
                return -1;
            }
            configured.set(dev_rd, static_cast<int>(smem));
        }
        const dim3 grid((y.max_sel + kRankRows - 1) / kRankRows, p.N);
        rank_decode_kernel<<<grid, kRankRows, smem, s>>>(a);
        POST_CHECK_LAUNCH("rank_decode_kernel");
        nms_offset_kernel<<<dim3((y.max_sel + 255) / 256, p.N), 256, 0, s>>>(a);
        POST_CHECK_LAUNCH("nms_offset_kernel");
        if (launches) *launches += 2;
    }
    if (p.nms_thresh > 0.f) {
        if (run_nms(nmsbox, sel_cnt, p.N, y.max_sel, p.nms_thresh, nms_scratch, y.nms_bytes, keep, nkeep, s, launches))
            return -1;
    } else {
        set_error("postprocess: nms_thresh <= 0 (NMS disabled) is not supported by the fused path");
        return -1;
    }
    {
        FinalArgs a;
        a.poly = poly;
        a.score = score;
        a.ctr = ctr;
        a.loc = loc;
        a.hbox = hbox;
        a.cls = cls;
        a.level = level;
        a.canon = canon;
        a.keep = keep;
        a.nkeep = nkeep;
        a.max_sel = y.max_sel;
        a.post_topk = p.post_nms_topk;
        a.do_postprocess = p.do_postprocess;
        a.sizes = p.sizes_dev;
        a.dets = p.dets;
        a.counts = p.counts;
        a.capacity = p.capacity;
        finalize_kernel<<<p.N, 1024, 0, s>>>(a);
        POST_CHECK_LAUNCH("finalize_kernel");
        if (launches) *launches += 1;
    }
    return 0;
}

int postprocess_debug_counts(const void* scratch, int N, int L, const int* level_hw, int num_classes, int pre_nms_topk,
                             int32_t* host_out, cudaStream_t s) {
    const Layout y = make_layout(N, L, level_hw, num_classes, pre_nms_topk);
    const uint8_t* base = static_cast<const uint8_t*>(scratch);
    cudaError_t e = cudaStreamSynchronize(s);
    for (int n = 0; n < N && e == cudaSuccess; ++n) {
        e = cudaMemcpy(host_out + 8 * n, base + y.o_cand_cnt + static_cast<size_t>(n) * 32, 5 * 4, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(host_out + 8 * n + 5, base + y.o_sel_cnt + static_cast<size_t>(n) * 4, 4, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(host_out + 8 * n + 6, base + y.o_nkeep + static_cast<size_t>(n) * 4, 4, cudaMemcpyDeviceToHost);
        host_out[8 * n + 7] = y.max_sel;
    }
    if (e != cudaSuccess) {
        set_error("postprocess_debug_counts: %s", cudaGetErrorString(e));
        return -1;
    }
    return 0;
}

int postprocess_debug_nms_stats(const void* scratch, int N, int L, const int* level_hw, int num_classes,
                                int pre_nms_topk, unsigned long long* host_out, cudaStream_t s) {
    const Layout y = make_layout(N, L, level_hw, num_classes, pre_nms_topk);
    return nms_read_stats(static_cast<const uint8_t*>(scratch) + y.o_nms, N, y.max_sel, host_out, s);
}

// ================================================================================================ stand-alone NMS hook
// ml_nms semantics for one image given boxes / scores / classes in arbitrary order.
__global__ void __launch_bounds__(1024) nms_prepare_kernel(const float* __restrict__ polys,
                                                           const float* __restrict__ scores,
                                                           const int* __restrict__ classes, int n, int vehicle_merge,
                                                           float* __restrict__ nmsbox, int* __restrict__ order,
                                                           int* __restrict__ count) {
    extern __shared__ unsigned long long s_keys[];
    __shared__ float s_min[32], s_max[32];
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    for (int i = threadIdx.x; i < np2; i += blockDim.x)
        s_keys[i] = i < n ? ((static_cast<unsigned long long>(__float_as_uint(scores[i])) << 32) |
                             (0xFFFFFFFFu - static_cast<unsigned>(i)))
                          : 0ull;
    __syncthreads();
    bitonic_sort_desc(s_keys, np2);
    float vmin = INFINITY, vmax = -INFINITY;
    for (int i = threadIdx.x; i < n * 8; i += blockDim.x) {
        vmin = fminf(vmin, polys[i]);
        vmax = fmaxf(vmax, polys[i]);
    }
    for (int o = 16; o > 0; o >>= 1) {
        vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
        vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    }
    if ((threadIdx.x & 31) == 0) {
        s_min[threadIdx.x >> 5] = vmin;
        s_max[threadIdx.x >> 5] = vmax;
    }
    __syncthreads();
    vmin = s_min[0];
    vmax = s_max[0];
    for (int w = 1; w < (blockDim.x >> 5); ++w) {
        vmin = fminf(vmin, s_min[w]);
        vmax = fmaxf(vmax, s_max[w]);
    }
    const float span = vmax - vmin + 1.0f;
    for (int r = threadIdx.x; r < n; r += blockDim.x) {
        const int src = static_cast<int>(0xFFFFFFFFu - static_cast<unsigned>(s_keys[r] & 0xFFFFFFFFu));
        order[r] = src;
        int c = classes ? classes[src] : 0;
        if (vehicle_merge && c == 5) c = 4;
        const float off = static_cast<float>(c) * span;
#pragma unroll
        for (int k = 0; k < 8; ++k) nmsbox[r * 8 + k] = polys[src * 8 + k] + off;
    }
    if (threadIdx.x == 0) *count = n;
}

__global__ void nms_gather_kernel(const int* __restrict__ order, const int* __restrict__ keep_pos,
                                  const int* __restrict__ nkeep, int* __restrict__ keep_out,
                                  int* __restrict__ nkeep_out) {
    const int nk = *nkeep;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nk; i += gridDim.x * blockDim.x)
        keep_out[i] = order[keep_pos[i]];
    if (blockIdx.x == 0 && threadIdx.x == 0) *nkeep_out = nk;
}

// ---- more than kMaxSorted boxes (the TTA union of up to 27 augmented copies x 1000 detections, tta.py:264-268):
// chunks of kMaxSorted keys are sorted in shared memory, merged pairwise by rank (all keys are distinct: score bits
// above, inverted input index below), and one CTA then derives the class offsets and writes the NMS boxes.
constexpr int kMaxNmsBoxes = 65536;

__global__ void __launch_bounds__(1024) nms_sort_chunk_kernel(const float* __restrict__ scores, int n,
                                                              unsigned long long* __restrict__ keys) {
    extern __shared__ unsigned long long s_keys[];
    const int base = blockIdx.x * kMaxSorted;
    const int cnt = min(kMaxSorted, n - base);
    int np2 = 1;
    while (np2 < cnt) np2 <<= 1;
    for (int i = threadIdx.x; i < np2; i += blockDim.x)
        s_keys[i] = i < cnt ? ((static_cast<unsigned long long>(__float_as_uint(scores[base + i])) << 32) |
                               (0xFFFFFFFFu - static_cast<unsigned>(base + i)))
                            : 0ull;
    __syncthreads();
    bitonic_sort_desc(s_keys, np2);
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) keys[base + i] = s_keys[i];
}

// merges descending runs [2r*len, 2r*len + len) and [2r*len + len, 2r*len + 2*len) of `in` (clipped to n) into `out`
__global__ void nms_merge_runs_kernel(const unsigned long long* __restrict__ in, unsigned long long* __restrict__ out,
                                      int n, int len) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int pair0 = (i / (2 * len)) * (2 * len);
    const int a0 = pair0, a1 = min(pair0 + len, n), b0 = a1, b1 = min(pair0 + 2 * len, n);
    const unsigned long long key = in[i];
    const bool in_a = i < a1;
    // number of elements of the OTHER run that come before `key` in descending order (keys are distinct)
    int lo = in_a ? b0 : a0, hi = in_a ? b1 : a1;
    const int other0 = lo;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (in[mid] > key)
            lo = mid + 1;
        else
            hi = mid;
    }
    out[pair0 + (i - (in_a ? a0 : b0)) + (lo - other0)] = key;
}

// the tail of nms_prepare_kernel for keys that are already sorted (one CTA)
__global__ void __launch_bounds__(1024) nms_prepare_sorted_kernel(const float* __restrict__ polys,
                                                                  const int* __restrict__ classes,
                                                                  const unsigned long long* __restrict__ keys, int n,
                                                                  int vehicle_merge, float* __restrict__ nmsbox,
                                                                  int* __restrict__ order, int* __restrict__ count) {
    __shared__ float s_min[32], s_max[32];
    float vmin = INFINITY, vmax = -INFINITY;
    for (int i = threadIdx.x; i < n * 8; i += blockDim.x) {
        vmin = fminf(vmin, polys[i]);
        vmax = fmaxf(vmax, polys[i]);
    }
    for (int o = 16; o > 0; o >>= 1) {
        vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
        vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    }
    if ((threadIdx.x & 31) == 0) {
        s_min[threadIdx.x >> 5] = vmin;
        s_max[threadIdx.x >> 5] = vmax;
    }
    __syncthreads();
    vmin = s_min[0];
    vmax = s_max[0];
    for (int w = 1; w < (blockDim.x >> 5); ++w) {
        vmin = fminf(vmin, s_min[w]);
        vmax = fmaxf(vmax, s_max[w]);
    }
    const float span = vmax - vmin + 1.0f;
    for (int r = threadIdx.x; r < n; r += blockDim.x) {
        const int src = static_cast<int>(0xFFFFFFFFu - static_cast<unsigned>(keys[r] & 0xFFFFFFFFu));
        order[r] = src;
        int c = classes ? classes[src] : 0;
        if (vehicle_merge && c == 5) c = 4;
        const float off = static_cast<float>(c) * span;
#pragma unroll
        for (int k = 0; k < 8; ++k) nmsbox[r * 8 + k] = polys[src * 8 + k] + off;
    }
    if (threadIdx.x == 0) *count = n;
}

size_t poly_nms_scratch_bytes(int n) {
    const size_t ms = n < 1 ? 1 : n;
    return a256(ms * 32) + a256(ms * 4) + a256(4) + a256(nms_scratch_bytes(1, static_cast<int>(ms))) + a256(ms * 4) +
           a256(4) + (n > kMaxSorted ? 2 * a256(ms * 8) : 0);
}

int launch_poly_nms(const float* polys, const float* scores, const int32_t* classes, int n, float thresh,
                    int vehicle_merge, int32_t* keep_out, int32_t* nkeep_out, void* scratch, size_t scratch_bytes,
                    cudaStream_t s) {
    if (n <= 0 || thresh <= 0.f) {
        // nms.py:22-25,63-64: nothing to suppress -> the reference returns the input unchanged / an empty index list
        cudaError_t e = cudaMemsetAsync(nkeep_out, 0, sizeof(int32_t), s);
        if (e != cudaSuccess) {
            set_error("poly_nms memset: %s", cudaGetErrorString(e));
            return -1;
        }
        if (n > 0) {
            set_error("poly_nms: nms_thresh <= 0 means 'no NMS' in the reference (ml_nms returns its input)");
            return -1;
        }
        return 0;
    }
    if (n > kMaxNmsBoxes) {
        set_error("poly_nms: n=%d exceeds the supported %d boxes per image", n, kMaxNmsBoxes);
        return -1;
    }
    if (poly_nms_scratch_bytes(n) > scratch_bytes) {
        set_error("poly_nms: scratch of %zu bytes is too small, need %zu", scratch_bytes, poly_nms_scratch_bytes(n));
        return -1;
    }
    const size_t ms = n;
    uint8_t* b = static_cast<uint8_t*>(scratch);
    float* nmsbox = reinterpret_cast<float*>(b);
    b += a256(ms * 32);
    int* order = reinterpret_cast<int*>(b);
    b += a256(ms * 4);
    int* count = reinterpret_cast<int*>(b);
    b += a256(4);
    void* nms_scratch = b;
    const size_t nms_bytes = nms_scratch_bytes(1, n);
    b += a256(nms_bytes);
    int* keep_pos = reinterpret_cast<int*>(b);
    b += a256(ms * 4);
    int* nk = reinterpret_cast<int*>(b);
    b += a256(4);
    if (n <= kMaxSorted) {
        int np2 = 1;
        while (np2 < n) np2 <<= 1;
        const size_t smem = static_cast<size_t>(np2) * 8;
        static DeviceOnce configured;
        int dev_np;
        if (static_cast<int>(smem) > configured.get(&dev_np)) {
            cudaError_t e = cudaFuncSetAttribute(nms_prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 static_cast<int>(smem));
            if (e != cudaSuccess) {
                set_error("nms_prepare_kernel smem attribute: %s", cudaGetErrorString(e));
                return -1;
            }
            configured.set(dev_np, static_cast<int>(smem));
        }
        nms_prepare_kernel<<<1, 1024, smem, s>>>(polys, scores, classes, n, vehicle_merge, nmsbox, order, count);
        POST_CHECK_LAUNCH("nms_prepare_kernel");
    } else {
        unsigned long long* keys_a = reinterpret_cast<unsigned long long*>(b);
        b += a256(ms * 8);
        unsigned long long* keys_b = reinterpret_cast<unsigned long long*>(b);
        const size_t smem = static_cast<size_t>(kMaxSorted) * 8;
        static DeviceOnce configured_chunk;
        int dev_chunk;
        if (!configured_chunk.get(&dev_chunk)) {
            cudaError_t e = cudaFuncSetAttribute(nms_sort_chunk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 static_cast<int>(smem));
            if (e != cudaSuccess) {
                set_error("nms_sort_chunk_kernel smem attribute: %s", cudaGetErrorString(e));
                return -1;
            }
            configured_chunk.set(dev_chunk, 1);
        }
        const int chunks = (n + kMaxSorted - 1) / kMaxSorted;
        nms_sort_chunk_kernel<<<chunks, 1024, smem, s>>>(scores, n, keys_a);
        POST_CHECK_LAUNCH("nms_sort_chunk_kernel");
        for (int len = kMaxSorted; len < n; len *= 2) {
            nms_merge_runs_kernel<<<(n + 255) / 256, 256, 0, s>>>(keys_a, keys_b, n, len);
            POST_CHECK_LAUNCH("nms_merge_runs_kernel");
            unsigned long long* t = keys_a;
            keys_a = keys_b;
            keys_b = t;
        }
        nms_prepare_sorted_kernel<<<1, 1024, 0, s>>>(polys, classes, keys_a, n, vehicle_merge, nmsbox, order, count);
        POST_CHECK_LAUNCH("nms_prepare_sorted_kernel");
    }
    if (run_nms(nmsbox, count, 1, n, thresh, nms_scratch, nms_bytes, keep_pos, nk, s, nullptr)) return -1;
    nms_gather_kernel<<<8, 256, 0, s>>>(order, keep_pos, nk, keep_out, nkeep_out);
    POST_CHECK_LAUNCH("nms_gather_kernel");
    return 0;
}

}  // namespace dafne
