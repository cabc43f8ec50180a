// Patch-merge polygon NMS in double precision (see merge_nms.cu).
#pragma once

namespace dafne {

// dets: host [offsets[nproblems]][9] doubles (8 coordinates + score), problem p = rows offsets[p] .. offsets[p+1]).
// keep_out (host, same length as dets rows): for problem p the kept LOCAL indices in descending score order at
// keep_out[offsets[p] .. offsets[p] + nkeep_out[p]). Synchronous; allocates and frees its own device memory.
int merge_nms_f64_batch_host(const double* dets, const int* offsets, int nproblems, double thresh, int device,
                             int* keep_out, int* nkeep_out);

// voc_eval's matching step: dets [nd][8] doubles (host, already sorted by confidence), det_image [nd] = index of each
// detection's image, gts [gt_offsets[nimages]][8] = ground truths of the class grouped by image. Writes per detection
// the best polygon IoU (-inf if no ground truth's horizontal box overlaps) and the LOCAL index of that ground truth (-1).
int voc_match_f64_host(const double* dets, const int* det_image, int nd, const double* gts, const int* gt_offsets,
                       int nimages, int device, double* ovmax_out, int* jmax_out);

}  // namespace dafne
