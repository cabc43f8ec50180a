// Patch-merge polygon NMS in double precision (see merge_nms.cu).
#pragma once

namespace dafne {

// dets: host [offsets[nproblems]][9] doubles (8 coordinates + score), problem p = rows offsets[p] .. offsets[p+1]).
// keep_out (host, same length as dets rows): for problem p the kept LOCAL indices in descending score order at
// keep_out[offsets[p] .. offsets[p] + nkeep_out[p]). Synchronous; allocates and frees its own device memory.
int merge_nms_f64_batch_host(const double* dets, const int* offsets, int nproblems, double thresh, int device,
                             int* keep_out, int* nkeep_out);

}  // namespace dafne
