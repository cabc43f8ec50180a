// ResNet stem (7x7 stride-2 pad-3 conv 3->64 + FrozenBN + ReLU + 3x3 stride-2 pad-1 max-pool; detectron2 v0.5 BasicStem
// via dafne/modeling/backbone/fpn.py:72) on the tensor cores, im2col-free, the pooling done in the epilogue.
//
// The normalised image is stored as fp16 NHWC4 (channel 3 = 0) on a zero canvas with 3 rows above/below and 4 pixels
// left/right of the image: [N][H+6][W+8][4]. For output pixel (oy, ox) and kernel row ky the 7 taps x 3 channels it
// needs are inside ONE contiguous 64-byte run of that canvas: padded row 2*oy + ky, padded pixels [2*ox, 2*ox + 8)
// (pixel 0 of the run and channel 3 get zero weights). A 5-D TMA tensor map whose pixel dimension advances by 16
// bytes while the innermost box is 64 bytes wide (overlapping windows) therefore delivers, per (tile, ky), a
// [128 pixels][32 elements] K-major operand tile straight into the SWIZZLE_64B layout tcgen05 reads:
//   GEMM  M = 128 conv pixels (16 x 8 patch), N = 64 channels, K = 7 k-blocks of 32 (two UMMA K=16 steps each).
//
// Max-pool fusion: the 16 x 8 conv patch starts at conv pixel (2*px0 - 1, 2*py0 - 1), so it contains every input of
// the 7 x 3 pooled pixels (px0.., py0..) -- 84 of its 128 conv pixels are pooled-output "payload", the rest is halo
// recomputed by the neighbouring tile (the conv is 6 % of a tensor-bound forward's FLOPs at most; what the fusion
// removes is the 128 B/pixel conv output going to HBM and coming back: 0.54 GB per 8 x 1024^2 batch). Conv pixels outside
// the conv output (the pool's padding) are replaced by 0, which equals the reference's -inf padding because every window
// holds at least one real post-ReLU value >= 0. fp16 rounding is monotonic, so max(round(x)) == round(max(x)): the
// result is bit-identical to rounding the conv output first and pooling it afterwards.
//
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-19 = four epilogue warpgroups.
// Warpgroup g drains accumulator stage g (tiles g, g+4, ... of this CTA): TMEM -> scale/shift/ReLU -> fp16 -> its own
// 16 KB staging tile in shared memory -> 3x3 max over the staged tile -> 16-byte global stores of the pooled pixels.
// One warpgroup would need ~2.5x the time the 14 MMAs of a tile take, hence four of them; with four, ncu's source page
// shows the epilogue warps waiting on the accumulator barrier (the MMA issue path is the critical one: see ptx.cuh,
// elect_one_sync).
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
// untouched. Two A loads per tile instead of seven, and the 28 KB of weights are loaded ONCE per CTA and stay resident
// (22 KB of overlapping-window operand per tile is what is left to fill). Measured on B200, R50 8 x 1024^2: unfused stem
// 130 us + pool 84 us -> fused 176 us -> resident weights 158 us -> elect.sync MMA issue 130 us (ncu r2n: 122 us, 41 %
// tensor pipe, 96 MB of DRAM traffic instead of 278 + 335 MB).
constexpr int kStages = 8;
constexpr int kEpiWgs = 4;          // epilogue warpgroups = accumulator stages
constexpr int kThreads = 128 + 128 * kEpiWgs;
constexpr int kABytes = 11 * 1024;  // 11 row pairs x 16 pixels x 32 fp16
constexpr int kBBytes = 64 * 64;    // 64 couts x 32 fp16, per kernel row
constexpr int kStageBytes = kABytes;
constexpr int kWBytes = 7 * kBBytes;  // all seven kernel rows, resident
constexpr int kEpiBytes = kEpiWgs * 16384;
constexpr int kAuxBytes = 256 + 2 * 64 * 4;
constexpr int kSmemBytes = kWBytes + kStages * kStageBytes + kEpiBytes + kAuxBytes;
constexpr int kTmemCols = 64 * kEpiWgs;
constexpr int kPoolW = 7, kPoolH = 3;  // pooled pixels per 16 x 8 conv patch
static_assert(kSmemBytes <= 232448, "stem: shared memory budget");
static_assert(kTmemCols == 256, "stem: TMEM allocation must be a power of two");

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

__device__ __forceinline__ uint32_t hmax2_bits(uint32_t a, uint32_t b) {
    uint32_t r;
    asm("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}

struct StemTile {
    int sx0, sy0, px0, py0, n;
};
__device__ __forceinline__ StemTile stem_tile(const StemParams& p, int t) {
    const int tx = t % p.tiles_x;
    const int r = t / p.tiles_x;
    StemTile c;
    c.px0 = tx * kPoolW;
    c.py0 = (r % p.tiles_y) * kPoolH;
    c.n = r / p.tiles_y;
    c.sx0 = 2 * c.px0 - 1;
    c.sy0 = 2 * c.py0 - 1;
    return c;
}
}  // namespace

__global__ void __launch_bounds__(kThreads, 1)
    stem_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ StemParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t s_w = smem_base;  // [7 ky][64 couts][32] fp16, SWIZZLE_64B
    const uint32_t s_tiles = smem_base + kWBytes;
    const uint32_t s_epi = s_tiles + kStages * kStageBytes;
    const uint32_t s_aux = s_epi + kEpiBytes;
    uint8_t* aux = smem + kWBytes + kStages * kStageBytes + kEpiBytes;
    const uint32_t bar_full = s_aux;                   // kStages x 8 B
    const uint32_t bar_empty = s_aux + 8 * kStages;    // kStages x 8 B
    const uint32_t bar_tfull = s_aux + 16 * kStages;   // kEpiWgs x 8 B
    const uint32_t bar_tempty = bar_tfull + 8 * kEpiWgs;
    const uint32_t bar_wfull = bar_tempty + 8 * kEpiWgs;
    volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(aux + 16 * kStages + 16 * kEpiWgs + 8);
    static_assert(16 * kStages + 16 * kEpiWgs + 12 <= 256, "stem: barrier header");
    float* s_scale = reinterpret_cast<float*>(aux + 256);
    float* s_shift = s_scale + 64;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(bar_full + 8 * i, 1);
            mbar_init(bar_empty + 8 * i, 1);
        }
        for (int i = 0; i < kEpiWgs; ++i) {
            mbar_init(bar_tfull + 8 * i, 1);
            mbar_init(bar_tempty + 8 * i, 4);
        }
        mbar_init(bar_wfull, 1);
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
        if (elect_one_sync()) {
            int stage = 0;
            uint32_t phase = 0;
            mbar_arrive_expect_tx(bar_wfull, kWBytes);
            for (int ky = 0; ky < 7; ++ky) tma_load_2d(s_w + ky * kBBytes, &tmB, bar_wfull, ky * 32, 0);
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                const StemTile tc = stem_tile(p, t);
                for (int par = 0; par < 2; ++par) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    const uint32_t full = bar_full + 8 * stage;
                    mbar_arrive_expect_tx(full, kABytes);
                    // conv column / row -1 (and those past the conv output) are outside the tensor map: zero-filled
                    // without touching memory; the epilogue discards those pixels anyway
                    tma_load_5d(s_tiles + stage * kStageBytes, &tmA, full, 0, tc.sx0, par, tc.sy0, tc.n);
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one_sync()) {
            constexpr uint32_t idesc = umma_idesc_f16(128, 64);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            mbar_wait(bar_wfull, 0);
            // One thread issues everything, and 14 small MMAs per tile leave it little time per instruction: both base
            // descriptors are built once, and with the loops unrolled every MMA's operands differ from them by
            // compile-time constants only (the address field of a descriptor is bits 0-13, in 16-byte units).
            const uint64_t a_base = umma_desc_sw64(s_tiles);
            const uint64_t b_base = umma_desc_sw64(s_w);
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
                const int acc = it % kEpiWgs;
                mbar_wait(bar_tempty + 8 * acc, ((it / kEpiWgs) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + acc * 64;
#pragma unroll
                for (int par = 0; par < 2; ++par) {
                    mbar_wait(bar_full + 8 * stage, phase);
                    tc_fence_after();
                    const uint64_t ad = a_base + static_cast<uint64_t>(stage * (kStageBytes >> 4));
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (j < 4 - par) {
                            // A: row pairs j .. j+7 of the box; B: kernel row 2j + par of the resident weights
#pragma unroll
                            for (int k = 0; k < 2; ++k)
                                umma_f16(d, ad + (j * 1024 >> 4) + 2 * k, b_base + ((2 * j + par) * kBBytes >> 4) + 2 * k,
                                         idesc, (par | j | k) != 0);
                        }
                    }
                    umma_commit(bar_empty + 8 * stage);
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(bar_tfull + 8 * acc);
            }
        }
    } else if (warp >= 4) {
        const int wg = (warp - 4) >> 2;
        const int wi = warp & 3;
        const int et = (threadIdx.x - 128) & 127;
        const int row = wi * 32 + lane;  // conv pixel of the patch: 16 wide, 8 high
        const int rx = row & 15, ry = row >> 4;
        const uint32_t buf = s_epi + wg * 16384;
        const uint32_t bar_id = 1 + wg;
        int it = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
            if (it % kEpiWgs != wg) continue;
            const StemTile tc = stem_tile(p, t);
            const int sx = tc.sx0 + rx, sy = tc.sy0 + ry;
            const bool inside = sx >= 0 && sx < p.Wo && sy >= 0 && sy < p.Ho;
            mbar_wait(bar_tfull + 8 * wg, (it / kEpiWgs) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(wi * 32) << 16) + wg * 64;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t v[32];
                DAFNE_TMEM_LD_X32(taddr + 32 * h, v);
                tmem_ld_wait();
                if (h == 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_tempty + 8 * wg);
                }
                uint32_t packed[16];
#pragma unroll
                for (int c = 0; c < 32; c += 4) {
                    const float4 sc = *reinterpret_cast<const float4*>(s_scale + 32 * h + c);
                    const float4 sh = *reinterpret_cast<const float4*>(s_shift + 32 * h + c);
                    const float a0 = fmaf(__uint_as_float(v[c]), sc.x, sh.x);
                    const float a1 = fmaf(__uint_as_float(v[c + 1]), sc.y, sh.y);
                    const float a2 = fmaf(__uint_as_float(v[c + 2]), sc.z, sh.z);
                    const float a3 = fmaf(__uint_as_float(v[c + 3]), sc.w, sh.w);
                    // ReLU on the conversion
                    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(packed[c >> 1]) : "f"(a1), "f"(a0));
                    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(packed[(c >> 1) + 1]) : "f"(a3), "f"(a2));
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t dst = buf + row * 128 + (((4 * h + q) ^ (row & 7)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst),
                                 "r"(inside ? packed[4 * q] : 0u), "r"(inside ? packed[4 * q + 1] : 0u),
                                 "r"(inside ? packed[4 * q + 2] : 0u), "r"(inside ? packed[4 * q + 3] : 0u)
                                 : "memory");
                }
            }
            named_bar_sync(bar_id, 128);
            // 21 pooled pixels x 8 channel vectors = 168 tasks for 128 threads
            for (int task = et; task < kPoolW * kPoolH * 8; task += 128) {
                const int q = task & 7, pp = task >> 3;
                const int ppx = pp % kPoolW, ppy = pp / kPoolW;
                const int px = tc.px0 + ppx, py = tc.py0 + ppy;
                if (px >= p.Wp || py >= p.Hp) continue;
                uint4 m = make_uint4(0, 0, 0, 0);  // post-ReLU values are >= 0
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        const int rr = (2 * ppy + dy) * 16 + 2 * ppx + dx;
                        const uint32_t src = buf + rr * 128 + ((q ^ (rr & 7)) << 4);
                        uint4 u;
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                     : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
                                     : "r"(src)
                                     : "memory");
                        m.x = hmax2_bits(m.x, u.x);
                        m.y = hmax2_bits(m.y, u.y);
                        m.z = hmax2_bits(m.z, u.z);
                        m.w = hmax2_bits(m.w, u.w);
                    }
                }
                __half* op = p.out + ((static_cast<size_t>(tc.n) * p.Hp + py) * p.Wp + px) * 64 + q * 8;
                *reinterpret_cast<uint4*>(op) = m;
            }
            named_bar_sync(bar_id, 128);  // the staging tile may be overwritten by this warpgroup's next tile
        }
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
    p.Ho = Ho;
    p.Wo = Wo;
    p.Hp = (Ho - 1) / 2 + 1;  // 3x3 stride 2 pad 1
    p.Wp = (Wo - 1) / 2 + 1;
    p.tiles_x = (p.Wp + kPoolW - 1) / kPoolW;
    p.tiles_y = (p.Hp + kPoolH - 1) / kPoolH;
    p.total_tiles = p.tiles_x * p.tiles_y * N;
    p.scale = scale;
    p.shift = shift;
    p.out = out;
    const cuuint64_t row_bytes = static_cast<cuuint64_t>(W + 8) * 8;
    {
        // {32 elements (64 B run), output x (16 B = 2 pixels per step), row parity, row pair, image}
        const cuuint64_t dims[5] = {32, (cuuint64_t)Wo, 2, (cuuint64_t)(Ho + 3), (cuuint64_t)N};
        const cuuint64_t str[4] = {16, row_bytes, 2 * row_bytes, (cuuint64_t)(H + 6) * row_bytes};
        const cuuint32_t box[5] = {32, 16, 1, 8 + 3, 1};  // 16 x 8 conv pixels, + 3 row pairs: all taps of a parity
        if (encode(&plan->tmA, canvas, 5, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B, "A")) return -1;
    }
    {
        const cuuint64_t dims[2] = {224, 64};
        const cuuint64_t str[1] = {224 * 2};
        const cuuint32_t box[2] = {32, 64};
        if (encode(&plan->tmB, w_packed, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B, "B")) return -1;
    }
    plan->grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
    return 0;
}

int stem_plan_launch(const StemPlan& pl, cudaStream_t s) {
    static DeviceOnce configured;
    int dev;
    if (!configured.get(&dev)) {
        cudaError_t e = cudaFuncSetAttribute(stem_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(stem_tc_kernel): %s", cudaGetErrorString(e));
            return -1;
        }
        configured.set(dev, 1);
    }
    stem_tc_kernel<<<pl.grid, kThreads, kSmemBytes, s>>>(pl.tmA, pl.tmB, pl.p);
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
