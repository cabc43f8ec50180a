// Tensor-core ResNet stem with the max-pool fused into its epilogue (see stem_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace dafne {

struct StemParams {
    int Ho, Wo;  // conv output (H/2 x W/2)
    int Hp, Wp;  // pooled output
    int tiles_x, tiles_y, total_tiles;
    const float* scale;  // folded FrozenBN, 64 channels
    const float* shift;
    __half* out;  // fp16 NHWC [N][Hp][Wp][64]
};

struct StemPlan {
    alignas(64) CUtensorMap tmA;
    alignas(64) CUtensorMap tmB;
    StemParams p;
    int grid;
};

// canvas: fp16 [N][H+6][W+8][4] (image at row 3, pixel 4; zeros elsewhere); w_packed: fp16 [64][7][8][4];
// out: fp16 NHWC [N][Hp][Wp][64], the max-pooled stem output (Hp = (H/2 - 1)/2 + 1).
int stem_plan_build(const __half* canvas, int N, int H, int W, const __half* w_packed, const float* scale,
                    const float* shift, __half* out, StemPlan* plan, int num_sms);
int stem_plan_launch(const StemPlan& plan, cudaStream_t stream);
int launch_pack_stem_weight_tc(const float* w_oihw, __half* out, cudaStream_t stream);

}  // namespace dafne
