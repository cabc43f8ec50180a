// Host-side description + plan for the tcgen05 implicit-GEMM convolution (see conv_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dafne {

// Function attributes (opt-in dynamic shared memory) and occupancy are PER DEVICE: a process that creates contexts on
// several GPUs must configure each kernel on each of them. `slot` is a function-local `static DeviceOnce`.
struct DeviceOnce {
    static constexpr int kMaxDevices = 64;
    int value[kMaxDevices] = {};  // 0 = not configured on that device yet; otherwise the cached value (>= 1)
    // current device's cached value, or 0 (then the caller configures and calls set())
    int get(int* dev_out) const {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
        *dev_out = dev;
        return value[dev];
    }
    void set(int dev, int v) { value[dev] = v; }
};

// What one convolution launch computes. Activations are NHWC fp16; weights are packed
// [Cout][tap][Cin] fp16 (tap = ky*k+kx); the epilogue is out = relu?(acc*scale + shift (+ residual)).
struct ConvDesc {
    const __half* in = nullptr;
    int N = 0, Hin = 0, Win = 0, Cin = 0;
    const __half* w = nullptr;
    int Cout = 0, ksize = 1, stride = 1;  // ksize 1|3 (pad = ksize/2), stride 1|2
    int Hout = 0, Wout = 0;               // filled by conv_out_dims()
    __half* out = nullptr;                // NHWC fp16 output (Cout % 64 == 0), or
    float* out_f32 = nullptr;             // NHWC fp32 output with row pitch out_ld (small-Cout prediction convs)
    int out_ld = 0;
    const float* scale = nullptr;  // per-Cout multiplier (folded FrozenBN), nullptr = 1
    const float* shift = nullptr;  // per-Cout addend (folded FrozenBN shift or conv bias), nullptr = 0
    int relu = 0;
    int reverse_m = 0;  // walk the pixel tiles from the last to the first (see ConvParams::reverse_m)
    // GroupNorm + ReLU of the INPUT applied while it is loaded (3x3 / stride 1 / Cout % 256 == 0 / Cin == 256, the tower
    // convolutions): `in` holds the previous layer's RAW conv output, in_gn_sums its statistics (the format gn_sums has),
    // in_gamma / in_beta the affine of the GroupNorm in between. nullptr: `in` is used as it is.
    const long long* in_gn_sums = nullptr;
    const float* in_gamma = nullptr;
    const float* in_beta = nullptr;
    const __half* residual = nullptr;  // NHWC fp16 [N, res_H, res_W, Cout]; added at (y>>res_shift, x>>res_shift)
    int res_H = 0, res_W = 0, res_shift = 0;
    // [N][Cout/8][2] 64-bit fixed point, accumulated: (sum * 2^20, sum of squares * 2^12) of the fp16-rounded output
    // over each image's pixels and each group of 8 channels. Fixed point makes the reduction order-independent.
    long long* gn_sums = nullptr;
};

constexpr float kGnSumScale = 1048576.0f;  // 2^20
constexpr float kGnSqScale = 4096.0f;      // 2^12

struct ConvParams {
    int N, Hout, Wout, Cin, Cout;
    int num_taps, cin_blocks;
    int tw, th, nb;
    int tiles_x, tiles_y, tiles_n, m_tiles, n_tiles, total_tiles;
    int tile_begin;  // index of this problem's first tile inside a grouped launch
    // Pixel tiles in DESCENDING order: for a convolution whose input was written by the launch right before it and is
    // larger than the 126 MB L2 -- conv1 of a bottleneck reading the previous block's 268 MB output -- the tail of that
    // tensor is what L2 still holds, so reading it back to front turns the first third of the reads into L2 hits.
    int reverse_m;
    int tap_view[9], tap_dy[9], tap_dx[9];
    const float* scale;
    const float* shift;
    int relu;
    const __half* residual;
    int res_H, res_W, res_shift;
    long long* gn_sums;
    float* out_f32;
    int out_ld;
    const long long* in_gn_sums;  // GroupNorm of the input applied on load (mode 5), or nullptr
    const float* in_gamma;
    const float* in_beta;
    const void* w_id;  // identity of the weight tensor (its device pointer): tiles with equal w_id share resident weights
};

// One convolution as the kernel sees it. A launch works through an array of these in device memory (tiles of all
// problems form one index space), so e.g. one head-tower layer over the five FPN levels is ONE persistent launch.
constexpr int kMaxConvProblems = 16;
struct alignas(128) ConvProblem {
    CUtensorMap tmA[4];
    CUtensorMap tmB;
    CUtensorMap tmOut;
    CUtensorMap tmRes;  // residual as 128 px x 64 ch boxes (same geometry as tmOut); valid iff the plan has res_tma
    ConvParams p;
};

struct ConvPlan {
    ConvProblem prob;  // host copy; conv_group_launch() needs it in device memory
    int block_n;
    int epi_wgs;  // epilogue warpgroups the shape wants (1 | 2, see ConvCfg)
    int mode;     // epilogue variant: 0 plain, 1 residual add, 2 GroupNorm statistics
    int res_tma;  // != 0: residual at the output's resolution, loaded by TMA into the epilogue ring; the value is the
                  // ring depth asked for (2..6)
    int row_shared;  // 3x3 stride 1, narrow N tile: 0 one A load per tap; 1 one per horizontal tap (18-row box);
                     // 2 one halo box for all nine taps; 3 the same with the weights resident in shared memory;
                     // 5 (256-wide tiles) halo boxes in their own two-slot ring, the weights of one tap per stage,
                     //   GroupNorm + ReLU of the input applied to the landed box by two transform warps
    int breg_bytes;  // mode 3: bytes of the resident weight region (9 x Cin/64 x BLOCK_N x 128)
    int grid;     // CTAs for a stand-alone launch
    double flops;  // 2*MACs, algorithmic (unpadded)
};

// fp16 tiled tensor map with 128-byte swizzle (shared with tail_tc.cu)
int encode_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
               const uint32_t* box, CUtensorMapL2promotion promo, const char* what);

inline void conv_out_dims(ConvDesc& d) {
    int pad = d.ksize / 2;
    d.Hout = (d.Hin + 2 * pad - d.ksize) / d.stride + 1;
    d.Wout = (d.Win + 2 * pad - d.ksize) / d.stride + 1;
}

// Returns 0 on success; on failure writes a message retrievable with dafne_last_error().
int conv_plan_build(const ConvDesc& d, ConvPlan* plan, int num_sms);
// Launch over `nprob` (<= kMaxConvProblems) problems with the same block_n, stored contiguously in DEVICE memory with
// p.tile_begin already assigned (prefix sums of p.total_tiles).
// mode, res_tma, row_shared and breg_bytes must be the same for every problem of the launch.
int conv_group_launch(const ConvProblem* dev_probs, int nprob, int total_tiles, int block_n, int epi_wgs, int mode,
                      int res_tma, int row_shared, int breg_bytes, int num_sms, cudaStream_t stream);

void set_error(const char* fmt, ...);
const char* get_error();

}  // namespace dafne
