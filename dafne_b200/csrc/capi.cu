// extern "C" surface of libdafne_b200.so (declared in include/dafne_b200.h). Pure glue: argument checks, error
// capture, and calls into the kernels' launchers. No C++ exception crosses this boundary.
#include "dafne_b200.h"

#include <cuda_runtime.h>

#include "conv_tc.cuh"
#include "elementwise.cuh"

using namespace dafne;

namespace {
int num_sms_cached() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
            sms = 148;
    }
    return sms;
}
}  // namespace

extern "C" {

const char* dafne_last_error(void) { return get_error(); }
int dafne_abi_version(void) { return DAFNE_ABI_VERSION; }

int dafne_conv_nhwc(const void* in, int N, int H, int W, int Cin, const void* w, int Cout, int ksize, int stride,
                    const float* scale, const float* shift, int relu, const void* residual, int res_H, int res_W,
                    int res_shift, float* gn_sums, void* out_f16, float* out_f32, int out_ld, void* stream) {
    if (!in || !w || ((out_f16 == nullptr) == (out_f32 == nullptr))) {
        set_error("dafne_conv_nhwc: need in, w and exactly one of out_f16 / out_f32");
        return -1;
    }
    ConvDesc d;
    d.in = static_cast<const __half*>(in);
    d.N = N;
    d.Hin = H;
    d.Win = W;
    d.Cin = Cin;
    d.w = static_cast<const __half*>(w);
    d.Cout = Cout;
    d.ksize = ksize;
    d.stride = stride;
    conv_out_dims(d);
    d.out = static_cast<__half*>(out_f16);
    d.out_f32 = out_f32;
    d.out_ld = out_ld;
    d.scale = scale;
    d.shift = shift;
    d.relu = relu;
    d.residual = static_cast<const __half*>(residual);
    d.res_H = res_H;
    d.res_W = res_W;
    d.res_shift = res_shift;
    d.gn_sums = gn_sums;
    ConvPlan plan;
    if (conv_plan_build(d, &plan, num_sms_cached())) return -1;
    return conv_plan_launch(plan, static_cast<cudaStream_t>(stream));
}

int dafne_gn_relu_nhwc(const void* in, void* out, int N, int HW, int C, int groups, const float* sums,
                       const float* gamma, const float* beta, float eps, void* stream) {
    return launch_gn_relu(static_cast<const __half*>(in), static_cast<__half*>(out), N, HW, C, groups, sums, gamma,
                          beta, eps, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
