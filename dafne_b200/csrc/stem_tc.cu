// ResNet stem (7x7 stride-2 pad-3 conv 3->64 + FrozenBN + ReLU; detectron2 v0.5 BasicStem via
// dafne/modeling/backbone/fpn.py:72) on the tensor cores, im2col-free.
//
// The normalised image is stored as fp16 NHWC4 (channel 3 = 0) on a zero canvas with 3 rows above/below and 4 pixels
// left/right of the image: [N][H+6][W+8][4]. For output pixel (oy, ox) and kernel row ky the 7 taps x 3 channels it
// needs are inside ONE contiguous 64-byte run of that canvas: padded row 2*oy + ky, padded pixels [2*ox, 2*ox + 8)
// (pixel 0 of the run and channel 3 get zero weights). A 5-D TMA tensor map whose pixel dimension advances by 16
// bytes while the innermost box is 64 bytes wide (overlapping windows) therefore delivers, per (tile, ky), a
// [128 pixels][32 elements] K-major operand tile straight into the SWIZZLE_64B layout tcgen05 reads:
//   GEMM  M = 128 output pixels (16 x 8 patch), N = 64 channels, K = 7 k-blocks of 32 (two UMMA K=16 steps each).
// Same warp roles as conv_tc.cu: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7 epilogue
// (TMEM -> scale/shift/ReLU -> fp16 -> swizzled smem -> TMA store). HBM-bound: 8 B/pixel in, 128 B/output pixel out.
#include <stdio.h>

#include "conv_tc.cuh"
#include "ptx.cuh"
#include "stem_tc.cuh"

namespace dafne {

namespace {
// One stage = one row PARITY of the 7 kernel rows: padded row 2*oy + ky = 2*(oy + ky/2) + (ky & 1), so the taps
// ky = 0, 2, 4, 6 read the even-row view from row-pair oy + 0..3 and ky = 1, 3, 5 the odd-row view from oy + 0..2.
// ONE box of th + 3 = 11 row pairs per parity therefore serves all of its taps: tap ky's operand is that box read
// from row ky/2 on -- the descriptor start moves by whole 1024-byte rows (two SWIZZLE_64B atoms), the layout is
// untouched. Two A loads per tile instead of seven (the kernel was bound by L2 -> shared-memory fills).
constexpr int kStages = 6;
constexpr int kABytes = 11 * 1024;  // 11 row pairs x 16 pixels x 32 fp16
constexpr int kBBytes = 64 * 64;    // 64 couts x 32 fp16, per kernel row
constexpr int kStageBytes = kABytes + 4 * kBBytes;
constexpr int kEpiBytes = 2 * 16384;
constexpr int kAuxBytes = 256 + 2 * 64 * 4;
constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + kAuxBytes;
constexpr int kTmemCols = 128;

// K-major SWIZZLE_64B operand tile: rows of 64 B, 8-row (512 B) swizzle atoms stacked along M/N.
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(512 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(4) << 61;
    return d;
}
}  // namespace

__global__ void __launch_bounds__(256, 1)
    stem_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ StemParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t s_tiles = smem_base;
    const uint32_t s_epi = smem_base + kStages * kStageBytes;
    const uint32_t s_aux = s_epi + kEpiBytes;
    uint8_t* aux = smem + kStages * kStageBytes + kEpiBytes;
    const uint32_t bar_full = s_aux;
    const uint32_t bar_empty = s_aux + 8 * kStages;
    const uint32_t bar_tfull = s_aux + 16 * kStages;
    const uint32_t bar_tempty = s_aux + 16 * kStages + 16;
    volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(aux + 16 * kStages + 32);
    float* s_scale = reinterpret_cast<float*>(aux + 256);
    float* s_shift = s_scale + 64;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmOut);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(bar_full + 8 * i, 1);
            mbar_init(bar_empty + 8 * i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_tfull + 8 * i, 1);
            mbar_init(bar_tempty + 8 * i, 4);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), kTmemCols);
        tmem_relinquish();
    }
    if (threadIdx.x >= 128 && threadIdx.x < 192) {
        const int c = threadIdx.x - 128;
        s_scale[c] = __ldg(p.scale + c);
        s_shift[c] = __ldg(p.shift + c);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                const int tx = t % p.tiles_x;
                const int r = t / p.tiles_x;
                const int ty = r % p.tiles_y, tn = r / p.tiles_y;
                const int x0 = tx * p.tw, y0 = ty * p.th, n0 = tn * p.nb;
                for (int par = 0; par < 2; ++par) {
                    const int nky = 4 - par;
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    const uint32_t full = bar_full + 8 * stage;
                    mbar_arrive_expect_tx(full, kABytes + nky * kBBytes);
                    const uint32_t sA = s_tiles + stage * kStageBytes;
                    tma_load_5d(sA, &tmA, full, 0, x0, par, y0, n0);
                    for (int j = 0; j < nky; ++j)
                        tma_load_2d(sA + kABytes + j * kBBytes, &tmB, full, (2 * j + par) * 32, 0);
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(128, 64);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + acc * 64;
                for (int par = 0; par < 2; ++par) {
                    mbar_wait(bar_full + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t sA = s_tiles + stage * kStageBytes;
                    for (int j = 0; j < 4 - par; ++j) {
                        const uint64_t ad = umma_desc_sw64(sA + j * 1024);  // row pairs j .. j+7 of the box
                        const uint64_t bd = umma_desc_sw64(sA + kABytes + j * kBBytes);
#pragma unroll
                        for (int k = 0; k < 2; ++k) umma_f16(d, ad + 2 * k, bd + 2 * k, idesc, (par | j | k) != 0);
                    }
                    umma_commit(bar_empty + 8 * stage);
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(bar_tfull + 8 * acc);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else if (warp >= 4) {
        const int wi = warp - 4;
        const int et = threadIdx.x - 128;
        const int row = wi * 32 + lane;
        int acc = 0;
        uint32_t acc_phase = 0;
        int store_buf = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
            const int tx = t % p.tiles_x;
            const int r = t / p.tiles_x;
            const int ty = r % p.tiles_y, tn = r / p.tiles_y;
            const int x0 = tx * p.tw, y0 = ty * p.th, n0 = tn * p.nb;
            mbar_wait(bar_tfull + 8 * acc, acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(wi * 32) << 16) + acc * 64;
            uint32_t v[64];
            DAFNE_TMEM_LD_X32(taddr, v);
            DAFNE_TMEM_LD_X32(taddr + 32, (v + 32));
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
            uint32_t packed[32];
#pragma unroll
            for (int c = 0; c < 64; c += 4) {
                const float4 sc = *reinterpret_cast<const float4*>(s_scale + c);
                const float4 sh = *reinterpret_cast<const float4*>(s_shift + c);
                const float a0 = fmaf(__uint_as_float(v[c]), sc.x, sh.x);
                const float a1 = fmaf(__uint_as_float(v[c + 1]), sc.y, sh.y);
                const float a2 = fmaf(__uint_as_float(v[c + 2]), sc.z, sh.z);
                const float a3 = fmaf(__uint_as_float(v[c + 3]), sc.w, sh.w);
                // ReLU on the conversion
                asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(packed[c >> 1]) : "f"(a1), "f"(a0));
                asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(packed[(c >> 1) + 1]) : "f"(a3), "f"(a2));
            }
            const uint32_t buf = s_epi + store_buf * 16384;
            if (et == 0) tma_store_wait_read<1>();
            named_bar_sync(1, 128);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const uint32_t dst = buf + row * 128 + ((q ^ (row & 7)) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(packed[4 * q]),
                             "r"(packed[4 * q + 1]), "r"(packed[4 * q + 2]), "r"(packed[4 * q + 3])
                             : "memory");
            }
            fence_proxy_async_smem();
            named_bar_sync(1, 128);
            if (et == 0) {
                tma_store_4d(&tmOut, buf, 0, x0, y0, n0);
                tma_store_commit();
            }
            store_buf ^= 1;
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
        if (et == 0) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int encode(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                  const cuuint32_t* box, CUtensorMapSwizzle swz, const char* what) {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !ptr) {
            set_error("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %s", cudaGetErrorString(e));
            return -1;
        }
        fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    }
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(stem %s) failed with CUresult %d", what, (int)r);
        return -1;
    }
    return 0;
}

int stem_plan_build(const __half* canvas, int N, int H, int W, const __half* w_packed, const float* scale,
                    const float* shift, __half* out, StemPlan* plan, int num_sms) {
    if (H % 2 || W % 2) {
        set_error("stem: H and W must be even (got %d x %d)", H, W);
        return -1;
    }
    const int Ho = H / 2, Wo = W / 2;
    StemParams& p = plan->p;
    p.tw = 16;
    p.th = 8;
    p.nb = 1;
    p.tiles_x = (Wo + p.tw - 1) / p.tw;
    p.tiles_y = (Ho + p.th - 1) / p.th;
    p.total_tiles = p.tiles_x * p.tiles_y * N;
    p.scale = scale;
    p.shift = shift;
    const cuuint64_t row_bytes = static_cast<cuuint64_t>(W + 8) * 8;
    {
        // {32 elements (64 B run), output x (16 B = 2 pixels per step), row parity, row pair, image}
        const cuuint64_t dims[5] = {32, (cuuint64_t)Wo, 2, (cuuint64_t)(Ho + 3), (cuuint64_t)N};
        const cuuint64_t str[4] = {16, row_bytes, 2 * row_bytes, (cuuint64_t)(H + 6) * row_bytes};
        const cuuint32_t box[5] = {32, (cuuint32_t)p.tw, 1, (cuuint32_t)p.th + 3, 1};  // + 3 row pairs: all taps of a parity
        if (encode(&plan->tmA, canvas, 5, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B, "A")) return -1;
    }
    {
        const cuuint64_t dims[2] = {224, 64};
        const cuuint64_t str[1] = {224 * 2};
        const cuuint32_t box[2] = {32, 64};
        if (encode(&plan->tmB, w_packed, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B, "B")) return -1;
    }
    {
        const cuuint64_t dims[4] = {64, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)N};
        const cuuint64_t str[3] = {128, (cuuint64_t)Wo * 128, (cuuint64_t)Ho * Wo * 128};
        const cuuint32_t box[4] = {64, (cuuint32_t)p.tw, (cuuint32_t)p.th, 1};
        if (encode(&plan->tmOut, out, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, "Out")) return -1;
    }
    plan->grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
    return 0;
}

int stem_plan_launch(const StemPlan& pl, cudaStream_t s) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(stem_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(stem_tc_kernel): %s", cudaGetErrorString(e));
            return -1;
        }
        configured = true;
    }
    stem_tc_kernel<<<pl.grid, 256, kSmemBytes, s>>>(pl.tmA, pl.tmB, pl.tmOut, pl.p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("stem_tc_kernel launch: %s", cudaGetErrorString(e));
        return -1;
    }
    return 0;
}

// fp32 [64, 3, 7, 7] -> fp16 [64][7 ky][8 px][4 ch]; px = kx + 1 (px 0 and ch 3 are zero)
__global__ void pack_stem_weight_tc_kernel(const float* __restrict__ w, __half* __restrict__ o) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 64 * 224) return;
    const int ch = i & 3, px = (i >> 2) & 7, ky = (i >> 5) % 7, co = i / 224;
    float v = 0.f;
    if (ch < 3 && px >= 1) v = w[((co * 3 + ch) * 7 + ky) * 7 + (px - 1)];
    o[i] = __float2half_rn(v);
}
int launch_pack_stem_weight_tc(const float* w, __half* out, cudaStream_t s) {
    pack_stem_weight_tc_kernel<<<(64 * 224 + 255) / 256, 256, 0, s>>>(w, out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("pack_stem_weight_tc_kernel launch: %s", cudaGetErrorString(e));
        return -1;
    }
    return 0;
}

}  // namespace dafne
