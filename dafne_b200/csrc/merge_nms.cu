// Polygon NMS of the reference's patch-merge step on device, in the reference's DOUBLE arithmetic.
//
//   py_cpu_nms_poly_fast(dets[n, 9] double, thresh)   dafne/utils/ResultMerge_multi_process.py:61-122
//     order by descending score; the best remaining box is kept and every other remaining box j is dropped unless
//     ovr(i, j) <= thresh, where ovr = polyiou.iou_poly(i, j) if the horizontal boxes overlap (hbb_ovr > 0) and
//     hbb_ovr (= 0) otherwise.
//   called once per (class file, full image) by nmsbynamedict / mergesingle (:155-221) on a 16-process host pool.
//
// Here: many (class, image) problems per call. The host sorts each problem by (score desc, index asc), and computes
// the horizontal boxes and their "+1" areas exactly as the reference does (same numpy-order double arithmetic); the
// device builds, for every 64 x 64 block of the upper triangle of every problem, the bitmask of "j is dropped by i"
// (merge_mask_kernel: hbox test, then iou_poly in double only where the reference would call it) and sweeps each
// problem's mask (merge_sweep_kernel). Compiled with -fmad=false -prec-div=true: bit-identical to the double oracle,
// which is pinned against the reference's own polyiou.cpp and against golden vectors produced by the reference's own
// py_cpu_nms_poly_fast.
#include <stdint.h>
#include <stdio.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "conv_tc.cuh"  // set_error
#include "merge_nms.cuh"
#include "polyiou_f64.cuh"

namespace dafne {

typedef unsigned long long u64;
constexpr int kRec = 13;  // per box: 8 coordinates, x1, y1, x2, y2, area

struct MergeTile {
    int row0, col0;      // absolute record index of the first row / column box of the tile
    int nrow, ncol;      // boxes in the tile (<= 64)
    int diag;            // 1: row block == column block (only j > i)
    long long mask_off;  // word index of mask[row0's problem-local row][column block]
    int nwords;          // mask words per row of this problem
};

__global__ void __launch_bounds__(64) merge_mask_kernel(const double* __restrict__ rec,
                                                        const MergeTile* __restrict__ tiles, double thresh,
                                                        u64* __restrict__ mask) {
    const MergeTile tl = tiles[blockIdx.x];
    __shared__ double col[64][kRec];
    const int t = threadIdx.x;
    if (t < tl.ncol) {
        const double* src = rec + static_cast<size_t>(tl.col0 + t) * kRec;
#pragma unroll
        for (int k = 0; k < kRec; ++k) col[t][k] = src[k];
    }
    __syncthreads();
    if (t >= tl.nrow) return;
    double r[kRec];
    const double* src = rec + static_cast<size_t>(tl.row0 + t) * kRec;
#pragma unroll
    for (int k = 0; k < kRec; ++k) r[k] = src[k];
    u64 bits = 0;
    for (int j = tl.diag ? t + 1 : 0; j < tl.ncol; ++j) {
        const double* c = col[j];
        // ResultMerge_multi_process.py:88-100 (no "+ 1" in w, h; "+ 1" areas)
        const double w = fmax(0.0, fmin(r[10], c[10]) - fmax(r[8], c[8]));
        const double h = fmax(0.0, fmin(r[11], c[11]) - fmax(r[9], c[9]));
        const double hbb_inter = w * h;
        double ovr = hbb_inter / (r[12] + c[12] - hbb_inter);
        if (ovr > 0.0) ovr = f64::iou_poly(r, c);  // :102-106
        if (!(ovr <= thresh)) bits |= 1ull << j;   // :117 keeps `ovr <= thresh`
    }
    mask[tl.mask_off + static_cast<long long>(t) * tl.nwords] = bits;
}

struct MergeProblem {
    int n, nwords;
    long long mask_off;  // first mask word of the problem
    int out_off;         // first slot of its keep list (= its first record index)
};

// one warp per problem: the sequential sweep over the sorted boxes
__global__ void __launch_bounds__(32) merge_sweep_kernel(const MergeProblem* __restrict__ probs,
                                                         const u64* __restrict__ mask, int* __restrict__ keep,
                                                         int* __restrict__ nkeep) {
    extern __shared__ u64 removed[];
    const MergeProblem pr = probs[blockIdx.x];
    const int lane = threadIdx.x;
    for (int w = lane; w < pr.nwords; w += 32) removed[w] = 0;
    __syncwarp();
    int nk = 0;
    for (int i = 0; i < pr.n; ++i) {
        const u64 word = removed[i >> 6];
        if ((word >> (i & 63)) & 1ull) continue;  // uniform: every lane reads the same word
        if (lane == 0) keep[pr.out_off + nk] = i;
        ++nk;
        const u64* row = mask + pr.mask_off + static_cast<long long>(i) * pr.nwords;
        for (int w = (i >> 6) + lane; w < pr.nwords; w += 32) removed[w] |= row[w];
        __syncwarp();
    }
    if (lane == 0) nkeep[blockIdx.x] = nk;
}

#define MERGE_CUDA(expr)                                                                   \
    do {                                                                                   \
        cudaError_t e__ = (expr);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            set_error("merge nms: %s: %s", #expr, cudaGetErrorString(e__));               \
            rc = -1;                                                                       \
            goto done;                                                                     \
        }                                                                                  \
    } while (0)

// Device scratch of the host-facing calls below: ONE grow-only allocation per device, kept for the life of the process
// and carved up per call -- in steady state these calls allocate nothing (the external poly_nms they stand in for does a
// cudaMalloc / cudaFree pair per buffer and call). Calls on the same device are serialised by the pool's mutex (they
// are synchronous anyway: host pointers in, host pointers out).
namespace {
struct HostCallPool {
    std::mutex mu;
    void* base[64] = {};
    size_t cap[64] = {};
    // returns nullptr (with the error set) on failure; `bytes` may be 0
    uint8_t* reserve(int device, size_t bytes) {
        if (device < 0 || device >= 64) {
            dafne::set_error("host call: device id %d out of range", device);
            return nullptr;
        }
        if (bytes > cap[device]) {
            if (base[device]) cudaFree(base[device]);
            base[device] = nullptr;
            cap[device] = 0;
            const size_t want = bytes + bytes / 4 + (1u << 20);
            cudaError_t e = cudaMalloc(&base[device], want);
            if (e != cudaSuccess) {
                dafne::set_error("host call: cudaMalloc(%zu): %s", want, cudaGetErrorString(e));
                return nullptr;
            }
            cap[device] = want;
        }
        return static_cast<uint8_t*>(base[device]);
    }
};
HostCallPool g_pool;
inline size_t up256(size_t v) { return (v + 255) / 256 * 256; }
}  // namespace


int merge_nms_f64_batch_host(const double* dets, const int* offsets, int nproblems, double thresh, int device,
                             int* keep_out, int* nkeep_out) {
    if (nproblems <= 0) return 0;
    if (!dets || !offsets || !keep_out || !nkeep_out) {
        set_error("merge nms: null argument");
        return -1;
    }
    const int total = offsets[nproblems];
    for (int p = 0; p < nproblems; ++p) {
        nkeep_out[p] = 0;
        if (offsets[p + 1] < offsets[p]) {
            set_error("merge nms: offsets must be non-decreasing");
            return -1;
        }
    }
    if (total == 0) return 0;
    // ---- host: order, horizontal boxes, areas (the reference's numpy arithmetic, in double)
    std::vector<double> rec(static_cast<size_t>(total) * kRec);
    std::vector<int> order(total);
    std::vector<MergeProblem> probs(nproblems);
    std::vector<MergeTile> tiles;
    long long mask_words = 0;
    int max_words = 0;
    for (int p = 0; p < nproblems; ++p) {
        const int b = offsets[p], n = offsets[p + 1] - b;
        int* ord = order.data() + b;
        for (int i = 0; i < n; ++i) ord[i] = i;
        std::stable_sort(ord, ord + n, [&](int x, int y) {
            return dets[static_cast<size_t>(b + x) * 9 + 8] > dets[static_cast<size_t>(b + y) * 9 + 8];
        });
        for (int i = 0; i < n; ++i) {
            const double* d = dets + static_cast<size_t>(b + ord[i]) * 9;
            double* r = rec.data() + static_cast<size_t>(b + i) * kRec;
            double x1 = d[0], y1 = d[1], x2 = d[0], y2 = d[1];
            for (int k = 0; k < 8; ++k) r[k] = d[k];
            for (int k = 1; k < 4; ++k) {
                x1 = std::min(x1, d[2 * k]);
                x2 = std::max(x2, d[2 * k]);
                y1 = std::min(y1, d[2 * k + 1]);
                y2 = std::max(y2, d[2 * k + 1]);
            }
            r[8] = x1;
            r[9] = y1;
            r[10] = x2;
            r[11] = y2;
            r[12] = (x2 - x1 + 1) * (y2 - y1 + 1);
        }
        MergeProblem& pr = probs[p];
        pr.n = n;
        pr.nwords = (n + 63) / 64;
        pr.mask_off = mask_words;
        pr.out_off = b;
        max_words = std::max(max_words, pr.nwords);
        for (int rb = 0; rb < pr.nwords; ++rb)
            for (int cb = rb; cb < pr.nwords; ++cb) {
                MergeTile t;
                t.row0 = b + rb * 64;
                t.col0 = b + cb * 64;
                t.nrow = std::min(64, n - rb * 64);
                t.ncol = std::min(64, n - cb * 64);
                t.diag = rb == cb;
                t.mask_off = mask_words + static_cast<long long>(rb) * 64 * pr.nwords + cb;
                t.nwords = pr.nwords;
                tiles.push_back(t);
            }
        mask_words += static_cast<long long>(n) * pr.nwords;
    }
    if (mask_words * 8 > (3ll << 30) || static_cast<size_t>(max_words) * 8 > 200 * 1024) {
        set_error("merge nms: %lld mask bytes / %d boxes in one problem exceed the supported size", mask_words * 8,
                  max_words * 64);
        return -1;
    }
    int rc = 0;
    double* d_rec = nullptr;
    MergeTile* d_tiles = nullptr;
    MergeProblem* d_probs = nullptr;
    u64* d_mask = nullptr;
    int *d_keep = nullptr, *d_nkeep = nullptr;
    std::vector<int> keep_sorted(total);
    std::lock_guard<std::mutex> lock(g_pool.mu);
    MERGE_CUDA(cudaSetDevice(device));
    {
        const size_t b_rec = up256(rec.size() * sizeof(double)), b_tiles = up256(tiles.size() * sizeof(MergeTile)),
                     b_probs = up256(probs.size() * sizeof(MergeProblem)),
                     b_mask = up256(static_cast<size_t>(std::max<long long>(mask_words, 1)) * 8),
                     b_keep = up256(static_cast<size_t>(total) * sizeof(int)),
                     b_nkeep = up256(static_cast<size_t>(nproblems) * sizeof(int));
        uint8_t* pool = g_pool.reserve(device, b_rec + b_tiles + b_probs + b_mask + b_keep + b_nkeep);
        if (!pool) return -1;
        d_rec = reinterpret_cast<double*>(pool);
        d_tiles = reinterpret_cast<MergeTile*>(pool + b_rec);
        d_probs = reinterpret_cast<MergeProblem*>(pool + b_rec + b_tiles);
        d_mask = reinterpret_cast<u64*>(pool + b_rec + b_tiles + b_probs);
        d_keep = reinterpret_cast<int*>(pool + b_rec + b_tiles + b_probs + b_mask);
        d_nkeep = reinterpret_cast<int*>(pool + b_rec + b_tiles + b_probs + b_mask + b_keep);
    }
    MERGE_CUDA(cudaMemcpy(d_rec, rec.data(), rec.size() * sizeof(double), cudaMemcpyHostToDevice));
    MERGE_CUDA(cudaMemcpy(d_tiles, tiles.data(), tiles.size() * sizeof(MergeTile), cudaMemcpyHostToDevice));
    MERGE_CUDA(cudaMemcpy(d_probs, probs.data(), probs.size() * sizeof(MergeProblem), cudaMemcpyHostToDevice));
    // rows of lower-triangle blocks are never written: the sweep only reads words w >= i / 64 of row i
    if (!tiles.empty()) {
        merge_mask_kernel<<<static_cast<unsigned>(tiles.size()), 64>>>(d_rec, d_tiles, thresh, d_mask);
        MERGE_CUDA(cudaGetLastError());
    }
    {
        const size_t smem = static_cast<size_t>(std::max(max_words, 1)) * 8;
        if (smem > 48 * 1024)
            MERGE_CUDA(cudaFuncSetAttribute(merge_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            static_cast<int>(smem)));
        merge_sweep_kernel<<<nproblems, 32, smem>>>(d_probs, d_mask, d_keep, d_nkeep);
        MERGE_CUDA(cudaGetLastError());
    }
    MERGE_CUDA(cudaMemcpy(keep_sorted.data(), d_keep, static_cast<size_t>(total) * sizeof(int), cudaMemcpyDeviceToHost));
    MERGE_CUDA(cudaMemcpy(nkeep_out, d_nkeep, static_cast<size_t>(nproblems) * sizeof(int), cudaMemcpyDeviceToHost));
    for (int p = 0; p < nproblems; ++p) {
        const int b = offsets[p];
        for (int k = 0; k < nkeep_out[p]; ++k) keep_out[b + k] = order[b + keep_sorted[b + k]];
    }
done:
    return rc;
}

}  // namespace dafne

// ================================================================================================ VOC matching (8f-4)
// The IoU-heavy inner loop of the reference's voc_eval (dafne/evaluation/voc_eval.py:133-186): for every detection the
// best polygon IoU against the ground truths of ITS image and the index of that ground truth. As in the reference the
// polygon IoU (double, iou_poly(GT, detection)) is evaluated only where the "+1" horizontal-box overlap is positive; the
// maximum keeps the FIRST of equal values (np.argmax). One thread per detection.
namespace dafne {

__global__ void voc_match_kernel(const double* __restrict__ dets, const int* __restrict__ det_image,
                                 const double* __restrict__ gts, const int* __restrict__ gt_offsets, int nd,
                                 double* __restrict__ ovmax_out, int* __restrict__ jmax_out) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= nd) return;
    double bb[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) bb[k] = dets[static_cast<size_t>(d) * 8 + k];
    double bx0 = bb[0], bx1 = bb[0], by0 = bb[1], by1 = bb[1];
#pragma unroll
    for (int k = 1; k < 4; ++k) {
        bx0 = fmin(bx0, bb[2 * k]);
        bx1 = fmax(bx1, bb[2 * k]);
        by0 = fmin(by0, bb[2 * k + 1]);
        by1 = fmax(by1, bb[2 * k + 1]);
    }
    const int img = det_image[d];
    double ovmax = -INFINITY;
    int jmax = -1;
    for (int j = gt_offsets[img]; j < gt_offsets[img + 1]; ++j) {
        double g[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) g[k] = gts[static_cast<size_t>(j) * 8 + k];
        double gx0 = g[0], gx1 = g[0], gy0 = g[1], gy1 = g[1];
#pragma unroll
        for (int k = 1; k < 4; ++k) {
            gx0 = fmin(gx0, g[2 * k]);
            gx1 = fmax(gx1, g[2 * k]);
            gy0 = fmin(gy0, g[2 * k + 1]);
            gy1 = fmax(gy1, g[2 * k + 1]);
        }
        // voc_eval.py:150-166
        const double iw = fmax(fmin(gx1, bx1) - fmax(gx0, bx0) + 1.0, 0.0);
        const double ih = fmax(fmin(gy1, by1) - fmax(gy0, by0) + 1.0, 0.0);
        const double inters = iw * ih;
        const double uni = (bx1 - bx0 + 1.0) * (by1 - by0 + 1.0) + (gx1 - gx0 + 1.0) * (gy1 - gy0 + 1.0) - inters;
        if (!(inters / uni > 0.0)) continue;
        const double ov = f64::iou_poly(g, bb);  // voc_eval.py:177-179: iou_poly(GT, bb)
        if (jmax < 0 || ov > ovmax) {            // np.max / np.argmax over the kept ones: first maximum
            ovmax = ov;
            jmax = j - gt_offsets[img];
        }
    }
    ovmax_out[d] = ovmax;
    jmax_out[d] = jmax;
}

int voc_match_f64_host(const double* dets, const int* det_image, int nd, const double* gts, const int* gt_offsets,
                       int nimages, int device, double* ovmax_out, int* jmax_out) {
    if (nd <= 0) return 0;
    if (!dets || !det_image || !gt_offsets || !ovmax_out || !jmax_out || nimages < 1) {
        set_error("voc match: bad arguments");
        return -1;
    }
    const int ng = gt_offsets[nimages];
    for (int i = 0; i < nd; ++i)
        if (det_image[i] < 0 || det_image[i] >= nimages) {
            set_error("voc match: detection %d refers to image %d of %d", i, det_image[i], nimages);
            return -1;
        }
    int rc = 0;
    double *d_dets = nullptr, *d_gts = nullptr, *d_ov = nullptr;
    int *d_img = nullptr, *d_off = nullptr, *d_j = nullptr;
    std::lock_guard<std::mutex> lock(g_pool.mu);
    MERGE_CUDA(cudaSetDevice(device));
    {
        const size_t b_dets = up256(static_cast<size_t>(nd) * 64), b_gts = up256(static_cast<size_t>(ng > 0 ? ng : 1) * 64),
                     b_ov = up256(static_cast<size_t>(nd) * 8), b_img = up256(static_cast<size_t>(nd) * 4),
                     b_off = up256(static_cast<size_t>(nimages + 1) * 4), b_j = up256(static_cast<size_t>(nd) * 4);
        uint8_t* pool = g_pool.reserve(device, b_dets + b_gts + b_ov + b_img + b_off + b_j);
        if (!pool) return -1;
        d_dets = reinterpret_cast<double*>(pool);
        d_gts = reinterpret_cast<double*>(pool + b_dets);
        d_ov = reinterpret_cast<double*>(pool + b_dets + b_gts);
        d_img = reinterpret_cast<int*>(pool + b_dets + b_gts + b_ov);
        d_off = reinterpret_cast<int*>(pool + b_dets + b_gts + b_ov + b_img);
        d_j = reinterpret_cast<int*>(pool + b_dets + b_gts + b_ov + b_img + b_off);
    }
    MERGE_CUDA(cudaMemcpy(d_dets, dets, static_cast<size_t>(nd) * 64, cudaMemcpyHostToDevice));
    if (ng > 0) MERGE_CUDA(cudaMemcpy(d_gts, gts, static_cast<size_t>(ng) * 64, cudaMemcpyHostToDevice));
    MERGE_CUDA(cudaMemcpy(d_img, det_image, static_cast<size_t>(nd) * 4, cudaMemcpyHostToDevice));
    MERGE_CUDA(cudaMemcpy(d_off, gt_offsets, static_cast<size_t>(nimages + 1) * 4, cudaMemcpyHostToDevice));
    voc_match_kernel<<<(nd + 63) / 64, 64>>>(d_dets, d_img, d_gts, d_off, nd, d_ov, d_j);
    MERGE_CUDA(cudaGetLastError());
    MERGE_CUDA(cudaMemcpy(ovmax_out, d_ov, static_cast<size_t>(nd) * 8, cudaMemcpyDeviceToHost));
    MERGE_CUDA(cudaMemcpy(jmax_out, d_j, static_cast<size_t>(nd) * 4, cudaMemcpyDeviceToHost));
done:
    return rc;
}

}  // namespace dafne
