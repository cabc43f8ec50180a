// HBM-bound helper kernels around the tcgen05 convolutions (see elementwise.cu).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dafne {

// NCHW uint8 / fp32 image batch -> (x - mean) / std, zero outside each image's (h, w), written as fp16 NHWC4
// (channel 3 = 0) onto the zero canvas [N][H+6][W+8][4] the tensor-core stem reads (image at row 3, pixel 4).
// sizes_dev: N rows of (h, w, out_h, out_w) int32 on the device.
int launch_preprocess(const void* images, int dtype, const int32_t* sizes_dev, int N, int H, int W, const float* mean3,
                      const float* std3, __half* out_canvas, cudaStream_t s);

// In-place y = relu(groupnorm(x)) of several NHWC fp16 tensors [N, HW, 256] (32 groups of 8 channels) in one launch.
constexpr int kMaxGnProblems = 16;
struct GnProblem {
    __half* x;              // normalised in place
    const long long* sums;  // [N][32][2] fixed-point statistics written by the producing conv (ConvDesc::gn_sums)
    const float* gamma;
    const float* beta;
    int N, HW;
};
struct GnGroupArgs {
    GnProblem prob[kMaxGnProblems];
    int block_begin[kMaxGnProblems + 1];
    int count;
    float eps;
};
int launch_gn_relu_group(const GnProblem* probs, int count, float eps, cudaStream_t s);

// y = relu(groupnorm(x)) from per-(image, group) fixed-point (sum, sumsq) (see ConvDesc::gn_sums); NHWC fp16,
// (C / groups) == 8.
int launch_gn_relu(const __half* in, __half* out, int N, int HW, int C, int groups, const long long* sums,
                   const float* gamma, const float* beta, float eps, cudaStream_t s);

// y = relu(x), fp16, n8 = number of 8-element vectors.
int launch_relu_copy(const __half* in, __half* out, size_t n8, cudaStream_t s);

// fp32 [Cout, Cin, k, k] -> fp16 [Cout][k*k][Cin]
int launch_pack_conv_weight(const float* w, int Cout, int Cin, int k, __half* out, cudaStream_t s);
// FrozenBN fold: scale = gamma * rsqrt(var + eps), shift = beta - mean * scale
int launch_fold_bn(const float* gamma, const float* beta, const float* mean, const float* var, float eps, int C,
                   float* scale, float* shift, cudaStream_t s);

}  // namespace dafne
