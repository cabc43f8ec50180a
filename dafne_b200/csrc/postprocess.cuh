// Rotated-box post-processing on device (see postprocess.cu): score/threshold/top-k, decode, corner sort,
// class-aware polygon NMS, post-NMS top-k, rescale/clip/filter. No host round trip anywhere.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace dafne {

struct PostLevel {
    const float* logits;  // [N, H*W, ld_logits], num_classes valid
    int ld_logits;
    const float* ctr;  // ctrness logit of (n, loc) at ctr[(n*HW + loc) * ld_ctr]
    int ld_ctr;
    const float* reg;  // 8 corner values of (n, loc) at reg[(n*HW + loc) * ld_reg + k]
    int ld_reg;
    const float* center;  // nullptr: `reg` already is corners_reg_pred. Else reg = (center[k%2] + reg[k]) * scale
    int ld_center;
    int H, W, stride;
};

struct PostParams {
    int N, L, num_classes;
    PostLevel lv[5];
    const float* scales_dev;  // [L] Scale parameters (dafne.py:47-53,409-411) or nullptr (= 1, only with center == nullptr)
    int sort_corners, thresh_with_ctr, pre_nms_topk, post_nms_topk, vehicle_merge, do_postprocess;
    float score_thresh, nms_thresh;
    const int32_t* sizes_dev;  // [N][4] = image h, w, output h, w
    float* dets;               // [N][capacity][20]
    int32_t* counts;           // [N]
    int capacity;
    void* scratch;
    size_t scratch_bytes;
};

size_t postprocess_scratch_bytes(int N, int L, const int* level_hw, int num_classes, int pre_nms_topk);
int launch_postprocess(const PostParams& p, cudaStream_t stream, int64_t* launches);

// Synchronous diagnostic: per image 8 ints = candidates above threshold per level (5), boxes entering NMS, boxes kept
// by NMS (before the post-NMS top-k), capacity of the per-image lists.
int postprocess_debug_counts(const void* scratch, int N, int L, const int* level_hw, int num_classes, int pre_nms_topk,
                             int32_t* host_out, cudaStream_t stream);

// Synchronous diagnostic: work counters of the NMS of the last launch_postprocess, host_out[8]: pairs consulted /
// pairs that needed the polygon clip for the diagonal panels [0], [1] and for the broadcast of kept rows [3], [4].
int postprocess_debug_nms_stats(const void* scratch, int N, int L, const int* level_hw, int num_classes,
                                int pre_nms_topk, unsigned long long* host_out, cudaStream_t stream);
int nms_read_stats(const void* nms_scratch, int N, int max_sel, unsigned long long* host_out, cudaStream_t stream);

int launch_sort_quadrilateral(const float* quads, float* out, int n, cudaStream_t stream);
int launch_poly_iou(const float* p, const float* q, float* iou, int n, cudaStream_t stream);
// fired[i] = 1 where the NMS pre-filter claims IoU(p[i], q[i]) == 0 without running the clip (test hook).
int launch_pair_filter(const float* p, const float* q, unsigned char* fired, int n, cudaStream_t stream);
int launch_term_filter(const float* p, const float* q, unsigned short* fired, unsigned short* nonzero, int n,
                       cudaStream_t stream);

// Lazily evaluated greedy polygon NMS of N images (nms.cu). nmsbox [N][max_sel][8] sorted by descending score with the
// class offsets applied, counts [N]; writes keep [N][max_sel] (kept positions, ascending) and nkeep [N].
size_t nms_scratch_bytes(int N, int max_sel);
int run_nms(const float* nmsbox, const int* counts, int N, int max_sel, float thr, void* scratch, size_t scratch_bytes,
            int* keep, int* nkeep, cudaStream_t stream, int64_t* launches);

size_t poly_nms_scratch_bytes(int n);
int launch_poly_nms(const float* polys, const float* scores, const int32_t* classes, int n, float thresh,
                    int vehicle_merge, int32_t* keep, int32_t* nkeep, void* scratch, size_t scratch_bytes,
                    cudaStream_t stream);

}  // namespace dafne
