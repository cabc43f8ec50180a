// Bottleneck tail: conv3 of block b and conv1 of block b + 1 as ONE two-GEMM tcgen05 kernel (see tail_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dafne {

// out  = relu(scale1 * (in  x W3^T) + shift1 + residual)      in  [N, H, W, K1],  W3 [N1][K1],  out, residual [N, H, W, N1]
// mid  = relu(scale2 * (out x W1^T) + shift2)                 W1 [N2][N1],        mid [N, H, W, N2]
// (detectron2 BottleneckBlock: conv3 + FrozenBN + shortcut + ReLU, then the next block's conv1 + FrozenBN + ReLU; both
// 1x1, stride 1.) `out` is rounded to fp16 before it feeds the second product, exactly as when it goes through HBM.
struct TailDesc {
    const __half* in = nullptr;
    int N = 0, H = 0, W = 0, K1 = 0;
    const __half* w3 = nullptr;
    int N1 = 0;
    const float *scale1 = nullptr, *shift1 = nullptr;
    const __half* residual = nullptr;
    __half* out = nullptr;
    const __half* w1 = nullptr;
    int N2 = 0;
    const float *scale2 = nullptr, *shift2 = nullptr;
    __half* mid = nullptr;
};

struct TailParams {
    int N, H, W, K1, N1, N2;
    int kb1;     // K1 / 64
    int chunks;  // N1 / 128
    int tw, th, nb, tiles_x, tiles_y, tiles_n, m_tiles;
    const float *scale1, *shift1, *scale2, *shift2;
    __half* mid;
};

struct alignas(128) TailProblem {
    CUtensorMap tmA;    // in:  box {64 ch, tw, th, nb}
    CUtensorMap tmB1;   // W3:  box {64 k, 128 n}
    CUtensorMap tmB2;   // W1:  box {64 k, min(N2, 128) n}
    CUtensorMap tmRes;  // residual, box {64 ch, tw, th, nb}
    CUtensorMap tmOut;  // out, same box
    TailParams p;
};

struct TailPlan {
    TailProblem prob;  // host copy; the kernel reads it from device memory
    int grid, smem_bytes, b_granules, o_slots;
    double flops;
};

// true if the pair of layers can run as one tail launch (shapes the kernel is built for)
bool tail_supported(int K1, int N1, int N2);
int tail_plan_build(const TailDesc& d, TailPlan* plan, int num_sms);
int tail_plan_launch(const TailProblem* dev_prob, const TailPlan& plan, cudaStream_t stream);

}  // namespace dafne
