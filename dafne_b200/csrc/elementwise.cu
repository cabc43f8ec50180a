// HBM-bound helper kernels of the dense forward: input normalisation, the 3-channel stem, max-pool, GroupNorm apply,
// and the one-time weight packing. Each restates a torch/ATen call the reference reaches on its inference path:
//   normalise + pad          dafne/modeling/one_stage_detector.py:100-107 (ImageList.from_tensors pads with 0 AFTER normalising)
//   stem conv+BN+ReLU, pool  detectron2 v0.5 BasicStem via dafne/modeling/backbone/fpn.py:72
//   GroupNorm(32) + ReLU     dafne/modeling/dafne/dafne.py:326-345
#include <stdint.h>

#include "conv_tc.cuh"
#include "elementwise.cuh"

namespace dafne {

#define DAFNE_CHECK_LAUNCH(name)                                        \
    do {                                                                \
        cudaError_t e__ = cudaGetLastError();                           \
        if (e__ != cudaSuccess) {                                       \
            set_error("%s launch: %s", name, cudaGetErrorString(e__)); \
            return -1;                                                  \
        }                                                               \
    } while (0)

// ------------------------------------------------------------------------------------------------ preprocess
// Output canvas: fp16 [N][H+6][W+8][4] with the image at row 3 / pixel 4 and zeros around it (stem_tc.cu reads
// 64-byte runs of it through TMA); every canvas position is written on every call.
// One thread = 4 consecutive canvas pixels (the canvas row pitch W + 8 and the image's left edge at pixel 4 are
// multiples of 4, so a group is either entirely left / right of the image columns or aligned with 4 image pixels):
// three 4-pixel loads (one per plane; 4 B for uint8, 16 B for float) and one 32-byte store per thread.
template <typename T>
struct alignas(4 * sizeof(T)) Px4 {
    T v[4];
};
template <typename T>
__global__ void preprocess_kernel(const T* __restrict__ img, const int32_t* __restrict__ sizes, int N, int H, int W,
                                  float m0, float m1, float m2, float s0, float s1, float s2,
                                  __half* __restrict__ out) {
    // grid (blocks over one padded image, N): 32-bit index arithmetic with ONE division per thread (the 64-bit
    // divisions of a flat grid-stride loop made this copy ALU-bound: 2.1 TB/s)
    const int PW = W + 8, PH = H + 6, PW4 = PW / 4;
    const unsigned per_image = static_cast<unsigned>(PH) * PW4;
    const float mean[3] = {m0, m1, m2}, sd[3] = {s0, s1, s2};
    const int n = blockIdx.y;
    for (unsigned j = blockIdx.x * blockDim.x + threadIdx.x; j < per_image; j += gridDim.x * blockDim.x) {
        const unsigned row = j / static_cast<unsigned>(PW4);
        const int x = static_cast<int>(j - row * PW4) * 4 - 4;
        const int y = static_cast<int>(row) - 3;
        const size_t i = static_cast<size_t>(n) * per_image + j;
        const int h = sizes[4 * n], w = sizes[4 * n + 1];  // rows of [h, w, out_h, out_w]
        float v[3][4];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int k = 0; k < 4; ++k) v[c][k] = 0.f;
        if (y >= 0 && y < h && x >= 0 && x < w) {
            const size_t plane = static_cast<size_t>(H) * W;
            const size_t base = static_cast<size_t>(n) * 3 * plane + static_cast<size_t>(y) * W + x;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const Px4<T> q = *reinterpret_cast<const Px4<T>*>(img + base + c * plane);  // x % 4 == 0, W % 32 == 0
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (x + k < w) v[c][k] = (static_cast<float>(q.v[k]) - mean[c]) / sd[c];
            }
        }
        uint4 o[2];
        uint32_t* ow = reinterpret_cast<uint32_t*>(o);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const __half2 a = __floats2half2_rn(v[0][k], v[1][k]);
            const __half2 b = __floats2half2_rn(v[2][k], 0.f);
            ow[2 * k] = *reinterpret_cast<const uint32_t*>(&a);
            ow[2 * k + 1] = *reinterpret_cast<const uint32_t*>(&b);
        }
        uint4* dst = reinterpret_cast<uint4*>(out) + i * 2;
        dst[0] = o[0];
        dst[1] = o[1];
    }
}

int launch_preprocess(const void* images, int dtype, const int32_t* sizes_dev, int N, int H, int W, const float* mean3,
                      const float* std3, __half* out, cudaStream_t s) {
    if (W % 4 != 0 || (reinterpret_cast<uintptr_t>(images) & 15) != 0) {
        set_error("preprocess: needs W %% 4 == 0 and 16-byte aligned images (W=%d)", W);
        return -1;
    }
    const size_t per_image = static_cast<size_t>(H + 6) * ((W + 8) / 4);
    if (per_image > 0x7fffffffull || N > 65535) {
        set_error("preprocess: image of %d x %d or batch of %d too large", H, W, N);
        return -1;
    }
    const int threads = 256;
    const unsigned bx = static_cast<unsigned>((per_image + threads - 1) / threads);
    const dim3 blocks(bx < 4096u ? bx : 4096u, static_cast<unsigned>(N));
    if (dtype == 0)
        preprocess_kernel<uint8_t><<<blocks, threads, 0, s>>>(static_cast<const uint8_t*>(images), sizes_dev, N, H, W,
                                                              mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2],
                                                              out);
    else if (dtype == 1)
        preprocess_kernel<float><<<blocks, threads, 0, s>>>(static_cast<const float*>(images), sizes_dev, N, H, W,
                                                            mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2],
                                                            out);
    else {
        set_error("preprocess: dtype %d (0 = uint8, 1 = float32)", dtype);
        return -1;
    }
    DAFNE_CHECK_LAUNCH("preprocess_kernel");
    return 0;
}

// ------------------------------------------------------------------------------------------------ fp16x2 max (ReLU copy)
// (the stem's 3x3 max-pool lives in the stem kernel's epilogue, stem_tc.cu)
__device__ __forceinline__ uint32_t hmax2_u32(uint32_t a, uint32_t b) {
    const __half2 r = __hmax2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
}

// ------------------------------------------------------------------------------------------------ GroupNorm + ReLU
// One launch normalises up to kMaxGnProblems tensors (e.g. one tower layer on all five FPN levels, both towers).
// Block = 256 threads = 32 channel vectors (8 channels = one group each) x 8 pixel lanes; it handles kGnPixPerBlock
// consecutive pixels of ONE image of one problem, so every thread keeps its group's affine y = x * a + b in registers
// and streams 16-byte vectors with four loads in flight.
constexpr int kGnPixPerBlock = 128;

__global__ void __launch_bounds__(256) gn_relu_group_kernel(const GnGroupArgs args) {
    int g = 0;
    while (g + 1 < args.count && static_cast<int>(blockIdx.x) >= args.block_begin[g + 1]) ++g;
    const GnProblem& pr = args.prob[g];
    const int blocks_per_img = (pr.HW + kGnPixPerBlock - 1) / kGnPixPerBlock;
    const int lb = blockIdx.x - args.block_begin[g];
    const int n = lb / blocks_per_img;
    const int p0 = (lb % blocks_per_img) * kGnPixPerBlock;
    const int cv = threadIdx.x & 31;  // channel vector = group index (C == 256, 32 groups of 8)
    const int pl = threadIdx.x >> 5;  // pixel lane 0..7
    const float inv_cnt = 1.0f / (static_cast<float>(pr.HW) * 8.0f);
    const float s1 = static_cast<float>(static_cast<double>(__ldg(pr.sums + (static_cast<size_t>(n) * 32 + cv) * 2)) *
                                        (1.0 / kGnSumScale));
    const float s2 = static_cast<float>(
        static_cast<double>(__ldg(pr.sums + (static_cast<size_t>(n) * 32 + cv) * 2 + 1)) * (1.0 / kGnSqScale));
    const float mean = s1 * inv_cnt;
    const float var = fmaxf(s2 * inv_cnt - mean * mean, 0.f);
    const float rstd = rsqrtf(var + args.eps);
    float ga[8], be[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        ga[k] = __ldg(pr.gamma + cv * 8 + k);
        be[k] = __ldg(pr.beta + cv * 8 + k);
    }
    uint4* base = reinterpret_cast<uint4*>(pr.x) + (static_cast<size_t>(n) * pr.HW) * 32 + cv;
    const int pend = min(p0 + kGnPixPerBlock, pr.HW);
    for (int q0 = p0 + pl; q0 < pend; q0 += 32) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int q = q0 + 8 * u;
            if (q < pend) v[u] = base[static_cast<size_t>(q) * 32];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int q = q0 + 8 * u;
            if (q >= pend) continue;
            const uint32_t vin[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
            uint32_t vo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&vin[j]));
                const float a0 = fmaxf((f.x - mean) * rstd * ga[2 * j] + be[2 * j], 0.f);
                const float a1 = fmaxf((f.y - mean) * rstd * ga[2 * j + 1] + be[2 * j + 1], 0.f);
                const __half2 h = __floats2half2_rn(a0, a1);
                vo[j] = *reinterpret_cast<const uint32_t*>(&h);
            }
            base[static_cast<size_t>(q) * 32] = make_uint4(vo[0], vo[1], vo[2], vo[3]);
        }
    }
}

int launch_gn_relu_group(const GnProblem* probs, int count, float eps, cudaStream_t s) {
    if (count < 1 || count > kMaxGnProblems) {
        set_error("gn_relu: %d tensors in one launch (1..%d supported)", count, kMaxGnProblems);
        return -1;
    }
    GnGroupArgs a;
    a.count = count;
    a.eps = eps;
    int total = 0;
    for (int i = 0; i < count; ++i) {
        a.prob[i] = probs[i];
        a.block_begin[i] = total;
        total += probs[i].N * ((probs[i].HW + kGnPixPerBlock - 1) / kGnPixPerBlock);
    }
    a.block_begin[count] = total;
    if (total == 0) return 0;
    gn_relu_group_kernel<<<total, 256, 0, s>>>(a);
    DAFNE_CHECK_LAUNCH("gn_relu_group_kernel");
    return 0;
}

int launch_gn_relu(const __half* in, __half* out, int N, int HW, int C, int groups, const long long* sums,
                   const float* gamma, const float* beta, float eps, cudaStream_t s) {
    if (C != 256 || groups != 32) {
        set_error("gn_relu: needs C == 256 and 32 groups (C=%d groups=%d)", C, groups);
        return -1;
    }
    if (in != out) {
        cudaError_t e = cudaMemcpyAsync(out, in, static_cast<size_t>(N) * HW * C * sizeof(__half),
                                        cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) {
            set_error("gn_relu copy: %s", cudaGetErrorString(e));
            return -1;
        }
    }
    GnProblem pr;
    pr.x = out;
    pr.sums = sums;
    pr.gamma = gamma;
    pr.beta = beta;
    pr.N = N;
    pr.HW = HW;
    return launch_gn_relu_group(&pr, 1, eps, s);
}

// ------------------------------------------------------------------------------------------------ relu copy
__global__ void relu_copy_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n8) {
    const __half2 z = __floats2half2_rn(0.f, 0.f);
    const uint32_t zu = *reinterpret_cast<const uint32_t*>(&z);
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n8;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        uint4 v = __ldg(in + i);
        v.x = hmax2_u32(v.x, zu);
        v.y = hmax2_u32(v.y, zu);
        v.z = hmax2_u32(v.z, zu);
        v.w = hmax2_u32(v.w, zu);
        out[i] = v;
    }
}

int launch_relu_copy(const __half* in, __half* out, size_t n8, cudaStream_t s) {
    if (n8 == 0) return 0;
    size_t blocks = (n8 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    relu_copy_kernel<<<static_cast<int>(blocks), 256, 0, s>>>(reinterpret_cast<const uint4*>(in),
                                                               reinterpret_cast<uint4*>(out), n8);
    DAFNE_CHECK_LAUNCH("relu_copy_kernel");
    return 0;
}

// ------------------------------------------------------------------------------------------------ weight packing
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, int Cout, int Cin, int k, __half* __restrict__ o) {
    const size_t total = static_cast<size_t>(Cout) * k * k * Cin;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int ci = i % Cin;
        const int tap = (i / Cin) % (k * k);
        const int co = i / (static_cast<size_t>(Cin) * k * k);
        o[i] = __float2half_rn(w[(static_cast<size_t>(co) * Cin + ci) * k * k + tap]);
    }
}
int launch_pack_conv_weight(const float* w, int Cout, int Cin, int k, __half* out, cudaStream_t s) {
    const size_t total = static_cast<size_t>(Cout) * k * k * Cin;
    size_t blocks = (total + 255) / 256;
    if (blocks > 4096) blocks = 4096;
    pack_conv_weight_kernel<<<static_cast<int>(blocks), 256, 0, s>>>(w, Cout, Cin, k, out);
    DAFNE_CHECK_LAUNCH("pack_conv_weight_kernel");
    return 0;
}

__global__ void fold_bn_kernel(const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                               int C, float* scale, float* shift) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C) return;
    const float sc = gamma[i] * rsqrtf(var[i] + eps);
    scale[i] = sc;
    shift[i] = beta[i] - mean[i] * sc;
}
int launch_fold_bn(const float* gamma, const float* beta, const float* mean, const float* var, float eps, int C,
                   float* scale, float* shift, cudaStream_t s) {
    fold_bn_kernel<<<(C + 255) / 256, 256, 0, s>>>(gamma, beta, mean, var, eps, C, scale, shift);
    DAFNE_CHECK_LAUNCH("fold_bn_kernel");
    return 0;
}

}  // namespace dafne
