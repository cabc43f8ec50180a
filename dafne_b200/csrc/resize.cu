// Bilinear resize of uint8 image planes on device, bit-identical to PIL.Image.resize(..., BILINEAR) -- what the
// reference's input side calls through detectron2's ResizeTransform (tools/plain_train_net.py:293-298,
// dafne/modeling/tta.py:76-93). Pillow's 8-bit resampling (src/libImaging/Resample.c), restated:
//   per output index xx of an axis: center = (xx + 0.5) * scale (scale = in / out), support = max(scale, 1),
//   xmin = max(0, (int)(center - support + 0.5)), xmax = min(in, (int)(center + support + 0.5)) - xmin,
//   weights = triangle((x + xmin - center + 0.5) / max(scale, 1)) normalised by their sum -- all in double --,
//   fixed point kk = (int)(0.5 + w * 2^22), out = clip8((2^21 + sum in[x + xmin] * kk[x]) >> 22);
//   horizontal pass into an 8-bit intermediate, then the vertical pass; a pass that keeps the size is skipped.
// Every thread recomputes the coefficients of its output column / row in double (compiled with -fmad=false: the same
// IEEE operations as the CPU), so no coefficient table has to be built on the host and copied.
#include <stdint.h>

#include "conv_tc.cuh"  // set_error
#include "resize.cuh"

namespace dafne {

constexpr int kPrecisionBits = 32 - 8 - 2;
constexpr int kMaxTaps = 64;  // 2 * ceil(max(scale, 1)) + 1 <= 64: down-scaling by up to ~31x

struct AxisCoef {
    int xmin, n;
    int kk[kMaxTaps];
};

__device__ __forceinline__ void axis_coefficients(int xx, int in_size, int out_size, AxisCoef& c) {
    const double scale = static_cast<double>(in_size) / static_cast<double>(out_size);
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = 1.0 * filterscale;
    const double ss = 1.0 / filterscale;
    const double center = 0.0 + (xx + 0.5) * scale;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double k[kMaxTaps];
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
        double v = (x + xmin - center + 0.5) * ss;
        if (v < 0.0) v = -v;
        const double w = v < 1.0 ? 1.0 - v : 0.0;
        k[x] = w;
        ww += w;
    }
    for (int x = 0; x < xmax; ++x) {
        if (ww != 0.0) k[x] /= ww;
        c.kk[x] = k[x] < 0 ? static_cast<int>(-0.5 + k[x] * (1 << kPrecisionBits))
                           : static_cast<int>(0.5 + k[x] * (1 << kPrecisionBits));
    }
    c.xmin = xmin;
    c.n = xmax;
}

__device__ __forceinline__ uint8_t clip8(int v) {
    v >>= kPrecisionBits;
    return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// in [P][H][W] -> out [P][H][nw]; block = one output column xx (coefficients shared through shared memory)
__global__ void __launch_bounds__(256) resize_h_kernel(const uint8_t* __restrict__ in, int P, int H, int W, int nw,
                                                       uint8_t* __restrict__ out) {
    __shared__ AxisCoef c;
    const int xx = blockIdx.x;
    if (threadIdx.x == 0) axis_coefficients(xx, W, nw, c);
    __syncthreads();
    const int rows = P * H;
    for (int r = blockIdx.y * blockDim.x + threadIdx.x; r < rows; r += gridDim.y * blockDim.x) {
        const uint8_t* src = in + static_cast<size_t>(r) * W + c.xmin;
        int acc = 1 << (kPrecisionBits - 1);
        for (int x = 0; x < c.n; ++x) acc += static_cast<int>(src[x]) * c.kk[x];
        out[static_cast<size_t>(r) * nw + xx] = clip8(acc);
    }
}

// in [P][H][W] -> out [P][nh][W]; block = one output row yy
__global__ void __launch_bounds__(256) resize_v_kernel(const uint8_t* __restrict__ in, int P, int H, int W, int nh,
                                                       uint8_t* __restrict__ out) {
    __shared__ AxisCoef c;
    const int yy = blockIdx.x;
    if (threadIdx.x == 0) axis_coefficients(yy, H, nh, c);
    __syncthreads();
    const int cols = P * W;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < cols; i += gridDim.y * blockDim.x) {
        const int p = i / W, x = i % W;
        const uint8_t* src = in + (static_cast<size_t>(p) * H + c.xmin) * W + x;
        int acc = 1 << (kPrecisionBits - 1);
        for (int y = 0; y < c.n; ++y) acc += static_cast<int>(src[static_cast<size_t>(y) * W]) * c.kk[y];
        out[(static_cast<size_t>(p) * nh + yy) * W + x] = clip8(acc);
    }
}

size_t resize_tmp_bytes(int planes, int H, int nw) { return static_cast<size_t>(planes) * H * nw; }

int launch_resize_bilinear_u8(const uint8_t* in, int planes, int H, int W, uint8_t* out, int nh, int nw, uint8_t* tmp,
                              cudaStream_t s) {
    if (planes < 1 || H < 1 || W < 1 || nh < 1 || nw < 1) {
        set_error("resize: bad sizes (%d planes, %dx%d -> %dx%d)", planes, H, W, nh, nw);
        return -1;
    }
    const double sx = static_cast<double>(W) / nw, sy = static_cast<double>(H) / nh;
    if (2 * static_cast<int>((sx < 1 ? 1 : sx) + 0.999999) + 1 > kMaxTaps ||
        2 * static_cast<int>((sy < 1 ? 1 : sy) + 0.999999) + 1 > kMaxTaps) {
        set_error("resize: down-scaling %dx%d -> %dx%d needs more than %d filter taps", H, W, nh, nw, kMaxTaps);
        return -1;
    }
    const bool need_h = nw != W, need_v = nh != H;
    if (!need_h && !need_v) {
        cudaError_t e = cudaMemcpyAsync(out, in, static_cast<size_t>(planes) * H * W, cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) {
            set_error("resize copy: %s", cudaGetErrorString(e));
            return -1;
        }
        return 0;
    }
    const uint8_t* src = in;
    if (need_h) {
        uint8_t* dst = need_v ? tmp : out;
        if (need_v && tmp == nullptr) {
            set_error("resize: a %zu-byte intermediate buffer is required", resize_tmp_bytes(planes, H, nw));
            return -1;
        }
        const int rows = planes * H;
        resize_h_kernel<<<dim3(nw, (rows + 2047) / 2048 < 1 ? 1 : (rows + 2047) / 2048), 256, 0, s>>>(src, planes, H, W,
                                                                                                    nw, dst);
        src = dst;
    }
    if (need_v) {
        const int cols = planes * nw;
        resize_v_kernel<<<dim3(nh, (cols + 2047) / 2048 < 1 ? 1 : (cols + 2047) / 2048), 256, 0, s>>>(src, planes, H, nw,
                                                                                                    nh, out);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("resize launch: %s", cudaGetErrorString(e));
        return -1;
    }
    return 0;
}

}  // namespace dafne
