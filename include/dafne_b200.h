/*
 * dafne_b200 -- C ABI of the B200-native DAFNe inference hot path (libdafne_b200.so).
 *
 * Scope: batched inference only -- dense forward (ResNet-50/101 + FPN + P6/P7 + DAFNe head) and the rotated-box
 * post-processing (threshold -> top-k -> decode -> corner sort -> polygon NMS -> top-1000 -> rescale/clip).
 * Plain C: pointers and sizes only, no torch / C++ types. Every function returns 0 on success and a negative
 * value on failure; dafne_last_error() then returns a thread-local message. No C++ exception crosses the ABI.
 * All `dev` pointers are CUDA device pointers on the context's device; `stream` is a cudaStream_t passed as void*
 * (NULL = legacy default stream). Unless stated otherwise no call synchronises the host with the device.
 *
 * Reference interfaces replaced (paths relative to the DAFNe reference tree):
 *   dafne/modeling/one_stage_detector.py:45-107   OneStageDetector.forward / preprocess_image  -> dafne_detect*
 *   dafne/modeling/backbone/fpn.py:16-91           ResNet + FPN + LastLevelP6P7                 -> dafne_forward_dense
 *   dafne/modeling/dafne/dafne.py:350-494          DAFNeHead.forward                            -> dafne_forward_dense
 *   dafne/modeling/dafne/dafne_outputs.py:733-925  predict_proposals / select_over_all_levels   -> dafne_postprocess
 *   dafne/utils/sort_corners.py:26-92              sort_quadrilateral                           -> dafne_sort_quadrilateral
 *   dafne/modeling/nms/nms.py:10-92                ml_nms / batched_nms_poly                    -> dafne_poly_nms
 *   dafne/modeling/nms/nms.py:91                   poly_gpu_nms(dets[n,9] host, thr, device_id) -> dafne_poly_nms_host
 *   tools/prepare_dota/polyiou.cpp:108-133         iou_poly                                     -> dafne_poly_iou
 *   dafne/utils/ResultMerge_multi_process.py:61-122  py_cpu_nms_poly_fast (patch merge, double) -> dafne_poly_nms_f64_host
 */
#ifndef DAFNE_B200_H
#define DAFNE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DAFNE_ABI_VERSION 1
#define DAFNE_MAX_LEVELS 5
/* One detection row = DAFNE_DET_STRIDE floats:
 *  [0..7] pred_corners x0,y0..x3,y3   [8..11] pred_boxes x1,y1,x2,y2   [12] score   [13] centerness
 *  [14] pred_class   [15] fpn_level   [16,17] location x,y   [18] canonical candidate index within the image
 *  (level-major, then location, then class)   [19] reserved (0) */
#define DAFNE_DET_STRIDE 20

typedef struct dafne_ctx dafne_ctx;

/* Flat model description; field meanings follow dafne/config/defaults.py:40-108 and the detectron2 keys the
 * pre-trained YAMLs dump (configs/pre-trained/dota-1.0_r101_ms.yaml:95-215). */
typedef struct dafne_model_spec {
    int32_t resnet_depth;      /* MODEL.RESNETS.DEPTH: 50 | 101 */
    int32_t num_classes;       /* MODEL.DAFNE.NUM_CLASSES (<= 32) */
    int32_t sort_corners;      /* MODEL.DAFNE.SORT_CORNERS */
    int32_t thresh_with_ctr;   /* MODEL.DAFNE.THRESH_WITH_CTR */
    int32_t pre_nms_topk;      /* MODEL.DAFNE.PRE_NMS_TOPK_TEST */
    int32_t post_nms_topk;     /* MODEL.DAFNE.POST_NMS_TOPK_TEST */
    float score_thresh;        /* MODEL.DAFNE.INFERENCE_TH_TEST */
    float nms_thresh;          /* MODEL.DAFNE.NMS_TH */
    int32_t num_levels;        /* len(MODEL.DAFNE.FPN_STRIDES) == 5 */
    int32_t fpn_strides[DAFNE_MAX_LEVELS];
    float pixel_mean[3];       /* MODEL.PIXEL_MEAN (applied in the input's channel order) */
    float pixel_std[3];        /* MODEL.PIXEL_STD */
    int32_t vehicle_merge;     /* 1 = reference behaviour: class 5 is treated as 4 inside NMS (nms.py:77-79) */
    int32_t reserved[7];
} dafne_model_spec;

const char* dafne_last_error(void);
int dafne_abi_version(void);

/* ------------------------------------------------------------------ context, weights, workspace */
int dafne_ctx_create(const dafne_model_spec* spec, int device, dafne_ctx** out);
void dafne_ctx_destroy(dafne_ctx* ctx);

/* Load weights by detectron2 state-dict name (e.g. "backbone.bottom_up.res2.0.conv1.weight",
 * "backbone.bottom_up.res2.0.conv1.norm.running_var", "proposal_generator.dafne_head.cls_tower.0.weight").
 * Tensors are fp32, contiguous, in the reference's native layouts (conv weights [Cout,Cin,kh,kw]) and live on the
 * device; the library packs them into its own fp16 layouts with CUDA kernels on `stream`. shapes = 4 int64 per
 * tensor (trailing dims 1). Unknown names are an error; names may arrive in any order and over several calls.
 * dafne_weights_finalize() checks that every tensor the model needs was provided and folds FrozenBN. */
int dafne_load_weights(dafne_ctx* ctx, int count, const char* const* names, const float* const* dev_ptrs,
                       const int64_t* shapes, void* stream);
int dafne_weights_finalize(dafne_ctx* ctx, void* stream);

/* The library never allocates per call: activations live in a caller-owned workspace, sized for a padded batch
 * N x 3 x H x W (H, W multiples of 32) and bound once per shape. Rebinding re-plans (tensor maps, launch table). */
int dafne_workspace_bytes(dafne_ctx* ctx, int N, int H, int W, size_t* bytes);
int dafne_bind_workspace(dafne_ctx* ctx, int N, int H, int W, void* dev_workspace, size_t bytes);

/* ------------------------------------------------------------------ the hot path */
/* images: NCHW, already padded to the bound N x 3 x H x W.  dtype 0 = uint8, 1 = float32 (un-normalised).
 * image_sizes: N x (h, w) int32 on the HOST = un-padded size of each image inside the padded batch. Pixels outside
 * (h, w) are replaced by 0 AFTER normalisation, as ImageList.from_tensors does.
 * Runs normalise -> ResNet -> FPN -> head; leaves logits / ctrness / corner regression in the workspace. */
int dafne_forward_dense(dafne_ctx* ctx, const void* dev_images, int dtype, const int32_t* image_sizes, void* stream);

/* Head outputs of the last dafne_forward_dense for one level, fp32 NHWC with row pitch *ld floats:
 * which 0 = class logits (num_classes valid), 1 = [ctrness logit, 8 corner deltas] (9 valid),
 * 2 = center regression (2 valid). *hw receives H_l, W_l. */
int dafne_head_output(dafne_ctx* ctx, int level, int which, const float** dev_ptr, int* ld, int* h, int* w);

/* Post-processing of head outputs held in the workspace.
 * image_sizes / output_sizes: N x (h, w) int32 on the host (output = the "height"/"width" the caller asked for).
 * The horizontal boxes are ALWAYS scaled to the output size, clipped and rows with an empty clipped box dropped
 * (detectron2 ProposalNetwork.forward -> detector_postprocess); do_postprocess gates only the rescale of the corners
 * and locations (OneStageDetector.forward(batched_inputs, do_postprocess), one_stage_detector.py:45-55, 78-98).
 * dev_dets: [N][capacity][DAFNE_DET_STRIDE] fp32, rows in descending score; dev_counts: [N] int32 = number of
 * detections the reference would return (may exceed `capacity` only through exact score ties at the top-k cut; rows
 * beyond capacity are dropped and the count still reports them). */
int dafne_postprocess(dafne_ctx* ctx, const int32_t* image_sizes, const int32_t* output_sizes, int do_postprocess,
                      float* dev_dets, int32_t* dev_counts, int capacity, void* stream);

/* Same post-processing on caller-provided head outputs in the reference's own form (fp32, NHWC views of the
 * reference's NCHW tensors): per level logits [N,H_l*W_l,C], reg [N,H_l*W_l,8] (= corners_reg_pred, i.e. already
 * (center.repeat + delta) * scale), ctr [N,H_l*W_l]. This is the bit-exact parity gate. */
int dafne_postprocess_external(dafne_ctx* ctx, int N, const int32_t* level_hw /* L x (H_l, W_l) */,
                               const float* const* dev_logits, const float* const* dev_reg,
                               const float* const* dev_ctr, const int32_t* image_sizes, const int32_t* output_sizes,
                               int do_postprocess, float* dev_dets, int32_t* dev_counts, int capacity,
                               void* dev_scratch, size_t scratch_bytes, void* stream);
int dafne_postprocess_scratch_bytes(dafne_ctx* ctx, int N, const int32_t* level_hw, size_t* bytes);

/* forward_dense + postprocess on device buffers. */
int dafne_detect(dafne_ctx* ctx, const void* dev_images, int dtype, const int32_t* image_sizes,
                 const int32_t* output_sizes, float* dev_dets, int32_t* dev_counts, int capacity, void* stream);

/* The same step as ONE instantiated CUDA graph (the launch list of a bound shape is static: dense forward, the fixed
 * sequence of post-processing / NMS panel launches). dafne_graph_capture runs the step once eagerly, then captures it
 * for exactly these buffers and sizes; dafne_graph_launch replays it on `stream` (new image CONTENTS in the same
 * dev_images buffer, results in the same dev_dets / dev_counts). Re-binding the workspace drops the graph. What it
 * buys is host time: one call instead of ~200 launches -- small batches (batch 1-3, TTA copies) are launch-bound. */
int dafne_graph_capture(dafne_ctx* ctx, const void* dev_images, int dtype, const int32_t* image_sizes,
                        const int32_t* output_sizes, int do_postprocess, float* dev_dets, int32_t* dev_counts,
                        int capacity, void* stream);
int dafne_graph_launch(dafne_ctx* ctx, void* stream);

/* The reference-facing call with HOST buffers: copies images H2D, runs dafne_detect, copies detections and counts
 * D2H and synchronises `stream`. host_images should be pinned for full copy bandwidth. */
int dafne_detect_host(dafne_ctx* ctx, const void* host_images, int dtype, const int32_t* image_sizes,
                      const int32_t* output_sizes, float* host_dets, int32_t* host_counts, int capacity,
                      void* stream);

/* Pipelined form of dafne_detect_host for serving loops (same arguments, same results). _begin enqueues the H2D copy
 * of this batch on the context's own copy stream into one of two staging buffers, then detection on `stream` and the
 * D2H copies of its results on a second copy stream (so the next batch's kernels never queue behind a copy), and
 * returns without synchronising; *ticket identifies the batch. _end blocks until that batch's
 * detections and counts are in the host buffers given to _begin. At most two batches are in flight: calling
 * _begin(i + 1) before _end(i) overlaps the H2D copy of batch i + 1 with the compute of batch i. The host buffers of
 * a batch must stay valid and untouched until its _end returns. */
int dafne_detect_host_begin(dafne_ctx* ctx, const void* host_images, int dtype, const int32_t* image_sizes,
                            const int32_t* output_sizes, float* host_dets, int32_t* host_counts, int capacity,
                            void* stream, int* ticket);
int dafne_detect_host_end(dafne_ctx* ctx, int ticket);

/* Device-side result record of a pipelined batch, for multi-GPU callers: [N][capacity][DAFNE_DET_STRIDE] fp32
 * detections immediately followed by [N] int32 counts, `*bytes` long, valid from dafne_detect_host_begin(ticket) (in
 * stream order) until the second dafne_detect_host_begin after it. One all-gather of this record on the compute stream
 * replaces the reference's pickled comm.gather (dafne/evaluation/dafne_evaluator.py:61-64) with no host bounce. */
int dafne_host_slot_wire(dafne_ctx* ctx, int ticket, const void** dev_wire, size_t* bytes, int* capacity);

/* Per-layer parity support. keep != 0 (set BEFORE dafne_bind_workspace) disables activation-memory reuse so that
 * every intermediate survives the forward; dafne_debug_activation then returns the NHWC fp16 tensor called `name`:
 * "pool" (stem conv + max-pool, one kernel), "res2.0" ... "res5.2", "p3" ... "p7", "cls_tower.l0" ... "corners_tower.l4". */
int dafne_debug_keep_activations(dafne_ctx* ctx, int keep);
int dafne_debug_activation(dafne_ctx* ctx, const char* name, const void** dev_ptr, int* N, int* H, int* W, int* C);

/* Diagnostic of the last dafne_postprocess / dafne_detect (synchronises the stream): host_out[N][8] = candidates above
 * the score threshold per level (5), boxes entering NMS, boxes kept by NMS before the post-NMS top-k, list capacity. */
int dafne_debug_post_counts(dafne_ctx* ctx, int32_t* host_out, void* stream);

/* Work counters of the rotated NMS of the last dafne_postprocess / dafne_detect (synchronises the stream), summed over
 * the batch: host_out[0], [1] = box pairs the greedy sweep consulted / pairs that needed the polygon clip (16 triangle
 * overlaps each) inside the 512-box diagonal panels; host_out[3], [4] = the same for the kept-rows x later-boxes
 * broadcast; the other slots are reserved (0). The reference's poly_gpu_nms clips n * (n - 1) / 2 pairs per image
 * (nms.py:91). */
int dafne_debug_nms_stats(dafne_ctx* ctx, uint64_t* host_out, void* stream);

/* Per-launch timing of the dense forward with CUDA events on the launching stream (bench.py's roofline numbers).
 * dafne_set_profiling(ctx, 1) makes every following dafne_forward_dense record an event after each launch;
 * dafne_get_profile (after the stream has been synchronised) returns, for launch i < *count: its duration in ms,
 * its algorithmic FLOPs (2*MACs) and HBM bytes, kind (0 = other, 1 = tcgen05 conv), the conv tile width block_n
 * and a label. Arrays may be NULL to query *count. */
typedef struct dafne_op_profile {
    float ms;
    int32_t kind, block_n, ksize, stride, cin, cout, hout, wout;
    double flops, bytes;
    char name[48];
} dafne_op_profile;
int dafne_set_profiling(dafne_ctx* ctx, int enable);
int dafne_get_profile(dafne_ctx* ctx, dafne_op_profile* ops, int capacity, int* count);

/* Counters for bench.py: kernels launched by this library since the last reset, and conv FLOPs (2*MACs). */
int dafne_stats(dafne_ctx* ctx, int64_t* kernel_launches, double* conv_flops, int reset);

/* ------------------------------------------------------------------ per-kernel hooks (tests, A/B) */
/* One convolution through the tcgen05 kernel. NHWC fp16 in/out, weights [Cout][k*k][Cin] fp16.
 * Exactly one of out_f16 / out_f32 is non-NULL (out_f32: Cout <= 32, row pitch out_ld).
 * dev_gn_sums (optional): [N][Cout/8][2] int64 fixed point, ACCUMULATED (zero it first): sum * 2^20 and sum of squares
 * * 2^12 of the fp16-rounded outputs per image and group of 8 channels. Integer accumulation makes GroupNorm
 * statistics bit-reproducible and independent of tiling and batch composition. */
int dafne_conv_nhwc(const void* dev_in_f16, int N, int H, int W, int Cin, const void* dev_w_f16, int Cout, int ksize,
                    int stride, const float* dev_scale, const float* dev_shift, int relu,
                    const void* dev_residual_f16, int res_H, int res_W, int res_shift, int64_t* dev_gn_sums,
                    void* dev_out_f16, float* dev_out_f32, int out_ld, void* stream);

/* Head-tower convolution with the PREVIOUS layer's GroupNorm + ReLU applied to its input while it is loaded (3x3,
 * stride 1, Cin = Cout = 256): dev_in_raw_f16 is the previous layer's raw convolution output, dev_in_gn_sums its
 * statistics (the dev_gn_sums format of dafne_conv_nhwc), in_gamma / in_beta the affine of the GroupNorm in between.
 *   out = conv3x3(relu(GroupNorm32(in_raw))) + shift,  dev_gn_sums += statistics of out (zero it first). */
int dafne_conv_gn_in_nhwc(const void* dev_in_raw_f16, int N, int H, int W, int Cin, const int64_t* dev_in_gn_sums,
                          const float* dev_in_gamma, const float* dev_in_beta, const void* dev_w_f16, int Cout,
                          const float* dev_shift, int64_t* dev_gn_sums, void* dev_out_f16, void* stream);

/* 1x1 / stride-1 convolution through the CTA-pair kernel (csrc/pair_tc.cu: tcgen05.mma.cta_group::2, 256 x 256 tiles over
 * two SMs, each loading half of the weight tile): detectron2 BottleneckBlock conv1 / conv3 + FrozenBN (+ shortcut) + ReLU.
 *   out = act(scale * (in x w^T) + shift (+ residual))   in [M, K], w [N, K], residual / out [M, N] fp16, M = N * H * W
 * K a multiple of 64, N a multiple of 256; residual may be NULL; act = ReLU when relu != 0. */
int dafne_conv1x1_pair_nhwc(const void* dev_in_f16, int64_t M, int K, const void* dev_w_f16, int N,
                            const float* dev_scale, const float* dev_shift, int relu, const void* dev_residual_f16,
                            void* dev_out_f16, void* stream);

/* Bottleneck tail through the two-GEMM tcgen05 kernel (csrc/tail_tc.cu): conv3 + FrozenBN + shortcut + ReLU of one
 * detectron2 BottleneckBlock and conv1 + FrozenBN + ReLU of the next one, both 1x1 / stride 1, in ONE launch:
 *   out = relu(scale1 * (in x w3^T) + shift1 + residual)   in [N,H,W,K1], w3 [N1][K1], residual / out [N,H,W,N1]
 *   mid = relu(scale2 * (out x w1^T) + shift2)              w1 [N2][N1], mid [N,H,W,N2]
 * K1 in {64,128,256}, N1 a multiple of 256, N2 in {64,128,256}; NHWC fp16 tensors, fp32 scale / shift. */
int dafne_bottleneck_tail_nhwc(const void* dev_in_f16, int N, int H, int W, int K1, const void* dev_w3_f16, int N1,
                               const float* dev_scale1, const float* dev_shift1, const void* dev_residual_f16,
                               void* dev_out_f16, const void* dev_w1_f16, int N2, const float* dev_scale2,
                               const float* dev_shift2, void* dev_mid_f16, void* stream);

/* GroupNorm apply + ReLU on NHWC fp16 from per-(image, group) sums produced by dafne_conv_nhwc. */
int dafne_gn_relu_nhwc(const void* dev_in_f16, void* dev_out_f16, int N, int HW, int C, int groups,
                       const int64_t* dev_gn_sums, const float* dev_gamma, const float* dev_beta, float eps,
                       void* stream);

/* sort_quadrilateral on device: quads [n,8] fp32 -> out [n,8] fp32 (sort_corners.py:26-92). */
int dafne_sort_quadrilateral(const float* dev_quads, float* dev_out, int n, void* stream);

/* Pairwise polygon IoU on device in the faithful fp32 arithmetic: iou[i] = IoU(p[i], q[i]). */
int dafne_poly_iou(const float* dev_p, const float* dev_q, float* dev_iou, int n, void* stream);

/* Test hook for the NMS pre-filter: fired[i] = 1 where the library skips the polygon clip of the pair (p[i] = the
 * higher-scored box, q[i] = the lower-scored one) because it can prove the faithful fp32 arithmetic yields IoU == 0.
 * Contract (tests/test_postprocess_gpu.py): fired[i] implies dafne_poly_iou gives exactly 0 for that pair. */
int dafne_poly_pair_filter(const float* dev_p, const float* dev_q, uint8_t* dev_fired, int n, void* stream);
/* Test hook for the per-TERM form of that filter: per pair, bit 4 * i + j of dev_fired = the filter declares the signed
 * overlap of edge triangle i of p with edge triangle j of q (polyiou.cpp:91-103) exactly zero; the same bit of
 * dev_nonzero = the faithful arithmetic's value of that term is not zero. Contract: (fired & nonzero) == 0. */
int dafne_poly_term_filter(const float* dev_p, const float* dev_q, uint16_t* dev_fired, uint16_t* dev_nonzero, int n,
                           void* stream);

/* Class-aware polygon NMS of one image (ml_nms -> batched_nms_poly -> poly_gpu_nms semantics): polys [n,8], scores
 * [n], classes [n] int32, all on device. dev_keep receives the kept input indices in descending score order
 * (ties: ascending input index), dev_nkeep their number. Workspace via dafne_poly_nms_scratch_bytes. */
int dafne_poly_nms(const float* dev_polys, const float* dev_scores, const int32_t* dev_classes, int n,
                   float nms_thresh, int vehicle_merge, int32_t* dev_keep, int32_t* dev_nkeep, void* dev_scratch,
                   size_t scratch_bytes, void* stream);
int dafne_poly_nms_scratch_bytes(int n, size_t* bytes);

/* Drop-in for the reference's native FFI  poly_gpu_nms(dets, thresh, device_id) / _poly_nms(keep_out, num_out,
 * polys_host, polys_num, polys_dim, thresh, device_id):  host pointers in and out, dets = [n][9] fp32 (8 coords +
 * score, class offsets already applied by the caller), synchronous. */
int dafne_poly_nms_host(int* keep_out, int* num_out, const float* polys_host, int polys_num, int polys_dim,
                        float nms_overlap_thresh, int device_id);

/* ------------------------------------------------------------------ next to the path: patch-merge NMS */
/* Device implementation of the reference's py_cpu_nms_poly_fast(dets, thresh)
 * (dafne/utils/ResultMerge_multi_process.py:61-122), the polygon NMS that merges per-patch detections back into full
 * DOTA images: DOUBLE precision, polyiou.iou_poly arithmetic (tools/prepare_dota/polyiou.cpp:108-133) evaluated only
 * for pairs whose horizontal boxes overlap, a box dropped unless ovr <= thresh. Host pointers in and out, synchronous.
 * dets_host = [n][9] doubles (8 coordinates + score); keep_out[n] receives the kept input indices in descending score
 * order (equal scores: ascending index), *num_out their number. */
int dafne_poly_nms_f64_host(const double* dets_host, int n, double thresh, int device_id, int32_t* keep_out,
                            int32_t* num_out);
/* The same for many (class, image) lists in one call -- what nmsbynamedict / mergesingle loop over on a 16-process
 * host pool (ResultMerge_multi_process.py:155-232): list p = rows offsets[p] .. offsets[p+1] of dets_host; its kept
 * LOCAL indices land at keep_out[offsets[p] .. offsets[p] + nkeep_out[p]). */
int dafne_poly_nms_f64_batch_host(const double* dets_host, const int32_t* offsets, int nproblems, double thresh,
                                  int device_id, int32_t* keep_out, int32_t* nkeep_out);

/* ------------------------------------------------------------------ next to the path: VOC AP with polygon IoU */
/* The matching step of the reference's voc_eval (dafne/evaluation/voc_eval.py:133-186) for one class: for every
 * detection (dets_host [nd][8] doubles, already sorted by descending confidence; det_image[d] = index of its image) the
 * best polygon IoU -- polyiou.iou_poly(GT, detection) in double, evaluated only where the "+1" horizontal-box overlap is
 * positive -- against the ground truths of that image (gts_host rows gt_offsets[i] .. gt_offsets[i+1]) and the local
 * index of that ground truth (first maximum); -inf / -1 when nothing overlaps. Host pointers, synchronous. The
 * sequential true/false-positive assignment and the AP integral stay with the caller (dafne_b200/voc_eval.py). */
int dafne_voc_match_f64_host(const double* dets_host, const int32_t* det_image, int nd, const double* gts_host,
                             const int32_t* gt_offsets, int nimages, int device_id, double* ovmax_out,
                             int32_t* jmax_out);

/* ------------------------------------------------------------------ next to the path: the input resize */
/* Bilinear resize of uint8 image planes on device, bit-identical to PIL.Image.resize((new_w, new_h), BILINEAR), which is
 * what detectron2's ResizeShortestEdge / ResizeTransform run on the host before the reference's model sees an image
 * (tools/plain_train_net.py:293-298, dafne/modeling/tta.py:76-93). dev_in [planes][h][w] (a CHW image has planes = 3),
 * dev_out [planes][new_h][new_w]; dev_tmp: planes * h * new_w bytes, needed when both sizes change (may be NULL
 * otherwise). No host synchronisation. */
int dafne_resize_bilinear_u8(const uint8_t* dev_in, int planes, int h, int w, uint8_t* dev_out, int new_h, int new_w,
                             uint8_t* dev_tmp, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DAFNE_B200_H */
