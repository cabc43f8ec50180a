// 1x1 / stride-1 convolution as a CTA-PAIR tcgen05 GEMM (cta_group::2, M = 256 per MMA): see pair_tc.cu.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dafne {

// out = act(scale * (in x W^T) + shift (+ residual))     in [M, K], W [N, K], residual / out [M, N], all fp16 row-major
// (NHWC tensors of a 1x1 / stride-1 convolution flattened over N * H * W); act = ReLU when relu != 0.
// detectron2 BottleneckBlock conv1 (+ FrozenBN + ReLU) and conv3 (+ FrozenBN + shortcut + ReLU) via
// dafne/modeling/backbone/fpn.py:72; the FPN lateral convolutions without a top-down add (fpn.py:16-37).
struct PairDesc {
    const __half* in = nullptr;
    long long M = 0;
    int K = 0;
    const __half* w = nullptr;
    int N = 0;
    const float *scale = nullptr, *shift = nullptr;
    int relu = 0;
    const __half* residual = nullptr;
    __half* out = nullptr;
    int reverse_m = 0;  // walk the pixel tiles back to front (the input's tail is what L2 still holds)
};

struct PairParams {
    int K, N, kbs, n_tiles, m_pairs, total;  // total = m_pairs * n_tiles pair tiles of 256 x 256
    int relu, reverse_m;
    const float *scale, *shift;
};

struct alignas(128) PairProblem {
    CUtensorMap tmA;    // in:  box {64 k, 128 rows}
    CUtensorMap tmB;    // W:   box {64 k, 128 rows}
    CUtensorMap tmRes;  // residual: box {64 ch, 128 rows}
    CUtensorMap tmOut;  // out: the same box
    PairParams p;
};

struct PairPlan {
    PairProblem prob;  // host copy; the kernel reads it from device memory
    int grid, smem_bytes, stages, slots, has_res;
    double flops, bytes;
};

// true if the layer can run on the pair kernel (K a multiple of 64, N a multiple of 256)
bool pair_supported(int K, int N);
int pair_plan_build(const PairDesc& d, PairPlan* plan, int num_sms);
int pair_plan_launch(const PairProblem* dev_prob, const PairPlan& plan, cudaStream_t stream);

}  // namespace dafne
