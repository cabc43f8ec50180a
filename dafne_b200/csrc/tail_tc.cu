// Bottleneck tail on tcgen05: conv3 of block b (1x1, FrozenBN, + shortcut, ReLU) and conv1 of block b + 1 (1x1, FrozenBN,
// ReLU) as ONE persistent two-GEMM kernel -- detectron2 v0.5 BottleneckBlock via dafne/modeling/backbone/fpn.py:72.
//
// Why: the 1x1 convolutions of res2-res4 sit below the B200 ridge (conv3 of res4: 119 FLOP/B). Run one after the other
// they move, per block pair of res4 at 32 x 1024^2, 64 + 256 + 256 MB (conv3: input, shortcut, output) and 256 + 64 MB
// (conv1: the same output read back, its own output): 896 MB. Fused, the block output is written once and never read
// back: 640 MB.
//
// One CTA owns a 128-pixel tile for ALL output channels:
//   GEMM 1  acc1[c] = in_tile[128 x K1] x W3[chunk c: 128 x K1]^T         per 128-channel chunk c of N1, two TMEM stages
//   epilogue  out_c = relu(scale1 * acc1 + shift1 + residual_c) -> fp16, staged in shared memory in the very K-major
//             128-byte-swizzled layout a TMA box load of that chunk would have (the same staging conv_tc.cu uses for
//             its TMA stores) -- so the staged chunk is at once the source of the TMA store of `out` and
//   GEMM 2  acc2 += out_c[128 x 128] x W1[N2 x chunk c]^T                  the A operand of the next block's conv1
//   final   mid = relu(scale2 * acc2 + shift2) -> fp16 -> global (64-byte vectors per pixel and 32 channels)
// The activation tile (K1 / 64 x 16 KB) is loaded once per pixel tile; weights stream through a ring of 16 KB granules
// (W3: one granule per 64-channel k-block of a chunk; W1: granules of <= 128 output rows, so every granule is a
// canonical operand tile); the residual chunk is TMA-loaded into the slot its output chunk is later staged in.
// TMEM: 2 x 128 columns for acc1, N2 <= 256 columns for acc2.
//
// Roles (384 threads): warp 0 = TMA producer (activations + weights), warp 1 = MMA issuer, warp 2 = TMEM allocator,
// warp 3 = residual producer, warps 4-7 / 8-11 = two epilogue warpgroups taking alternate chunks (acc1 stage w belongs
// to warpgroup w) and half of acc2's columns each.
#include <stdio.h>

#include "conv_tc.cuh"  // set_error, encode_map, DeviceOnce
#include "ptx.cuh"
#include "tail_tc.cuh"

namespace dafne {

namespace {
constexpr int kGran = 16384;      // 128 rows x 64 channels fp16: one operand tile
constexpr int kSlot = 2 * kGran;  // one output / residual chunk: 128 pixels x 128 channels
constexpr int kTailThreads = 384;
constexpr int kMaxGran = 8, kMaxSlots = 4;
constexpr int kAuxBars = 512;                      // barrier block + TMEM pointer
constexpr int kAuxBytes = kAuxBars + 2 * 128 * 8;  // + one (scale, shift) table of 128 channels per warpgroup
constexpr int kTailSmemMax = 232448;               // 227 KB

// barrier offsets inside the aux block
constexpr int oAFull = 0, oAEmpty = 8, oBFull = 16, oBEmpty = 16 + 8 * kMaxGran, oT1Full = 144, oT1Empty = 160,
              oRFull = 176, oOReady = 176 + 8 * kMaxSlots, oOFree = 176 + 16 * kMaxSlots, oT2Full = 272, oT2Empty = 280,
              oTmemPtr = 288;

struct TileXY {
    int x0, y0, n0;
};
__device__ __forceinline__ TileXY tail_tile(const TailParams& p, int t) {
    TileXY c;
    const int tx = t % p.tiles_x;
    const int r = t / p.tiles_x;
    c.x0 = tx * p.tw;
    c.y0 = (r % p.tiles_y) * p.th;
    c.n0 = (r / p.tiles_y) * p.nb;
    return c;
}

// relu(acc * scale + shift (+ residual)) for 32 channels -> 16 packed fp16 pairs
template <bool HAS_RES>
__device__ __forceinline__ void tail_math32(const uint32_t (&v)[32], const float2* tab, const uint4 (&res)[4],
                                            uint32_t (&packed)[16]) {
    const uint32_t* rw = reinterpret_cast<const uint32_t*>(res);
#pragma unroll
    for (int c = 0; c < 32; c += 2) {
        const float4 tb = *reinterpret_cast<const float4*>(&tab[c]);
        float a0 = fmaf(__uint_as_float(v[c]), tb.x, tb.y);
        float a1 = fmaf(__uint_as_float(v[c + 1]), tb.z, tb.w);
        if (HAS_RES) {
            const float2 rf = __half22float2(*reinterpret_cast<const __half2*>(&rw[c >> 1]));
            a0 += rf.x;
            a1 += rf.y;
        }
        uint32_t h;
        asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(a1), "f"(a0));
        packed[c >> 1] = h;
    }
}
}  // namespace

__global__ void __launch_bounds__(kTailThreads, 1)
    tail_tc_kernel(const TailProblem* __restrict__ prob, int nb_gran, int n_slots) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const TailParams& p = prob->p;
    const int kb1 = p.kb1, chunks = p.chunks, N2 = p.N2;
    const int n2rows = N2 < 128 ? N2 : 128;  // output rows of one W1 granule
    const int n2h = N2 / n2rows;             // granules per 64-channel k-block of GEMM 2
    const int m_tiles = p.m_tiles;

    const uint32_t smem_base = smem_u32(smem);
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;
    // carve-up: [activation tile: kb1 granules] [weight ring: nb_gran granules] [chunk slots: n_slots x 32 KB] [aux]
    const uint32_t sA = smem_base;
    const uint32_t sB = sA + kb1 * kGran;
    const uint32_t sO = sB + nb_gran * kGran;
    const uint32_t s_aux = sO + n_slots * kSlot;
    uint8_t* aux = smem + (kb1 + nb_gran) * kGran + n_slots * kSlot;
    volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(aux + oTmemPtr);
    float2* tables = reinterpret_cast<float2*>(aux + kAuxBars);

    if (threadIdx.x == 0 && (smem_base & 1023u) != 0) {
        printf("dafne tail_tc: dynamic smem base not 1024-aligned (%u)\n", smem_base);
        __trap();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&prob->tmA);
        tma_prefetch_desc(&prob->tmB1);
        tma_prefetch_desc(&prob->tmB2);
        tma_prefetch_desc(&prob->tmRes);
        tma_prefetch_desc(&prob->tmOut);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(s_aux + oAFull, 1);
        mbar_init(s_aux + oAEmpty, 1);
        for (int i = 0; i < kMaxGran; ++i) {
            mbar_init(s_aux + oBFull + 8 * i, 1);
            mbar_init(s_aux + oBEmpty + 8 * i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(s_aux + oT1Full + 8 * i, 1);
            mbar_init(s_aux + oT1Empty + 8 * i, 4);  // one arrive per warp of the owning warpgroup
        }
        for (int i = 0; i < kMaxSlots; ++i) {
            mbar_init(s_aux + oRFull + 8 * i, 1);
            mbar_init(s_aux + oOReady + 8 * i, 1);
            mbar_init(s_aux + oOFree + 8 * i, 2);  // GEMM 2 has consumed the slot AND its TMA store has read it
        }
        mbar_init(s_aux + oT2Full, 1);
        mbar_init(s_aux + oT2Empty, 8);  // every epilogue warp
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer: activation tile + weight granules
        if (elect_one_sync()) {
            int bs = 0;
            uint32_t bph = 0;
            auto next_gran = [&](uint32_t bytes) -> uint32_t {
                mbar_wait(s_aux + oBEmpty + 8 * bs, bph ^ 1);
                mbar_arrive_expect_tx(s_aux + oBFull + 8 * bs, bytes);
                return sB + bs * kGran;
            };
            auto advance = [&]() {
                if (++bs == nb_gran) {
                    bs = 0;
                    bph ^= 1;
                }
            };
            auto load_b2 = [&](int cc) {
                for (int h = 0; h < 2; ++h)
                    for (int nh = 0; nh < n2h; ++nh) {
                        const uint32_t full = s_aux + oBFull + 8 * bs;
                        const uint32_t dst = next_gran(n2rows * 128);
                        tma_load_2d(dst, &prob->tmB2, full, cc * 128 + h * 64, nh * n2rows);
                        advance();
                    }
            };
            int it = 0;
            for (int t = blockIdx.x; t < m_tiles; t += gridDim.x, ++it) {
                const TileXY tc = tail_tile(p, t);
                mbar_wait(s_aux + oAEmpty, (it & 1) ^ 1);  // GEMM 1 of the previous tile has read the old tile
                mbar_arrive_expect_tx(s_aux + oAFull, kb1 * kGran);
                for (int kb = 0; kb < kb1; ++kb)
                    tma_load_4d(sA + kb * kGran, &prob->tmA, s_aux + oAFull, kb * 64, tc.x0, tc.y0, tc.n0);
                for (int c = 0; c < chunks; ++c) {
                    for (int kb = 0; kb < kb1; ++kb) {
                        const uint32_t full = s_aux + oBFull + 8 * bs;
                        const uint32_t dst = next_gran(kGran);
                        tma_load_2d(dst, &prob->tmB1, full, kb * 64, c * 128);
                        advance();
                    }
                    if (c >= 1) load_b2(c - 1);  // the MMA warp consumes in exactly this order
                }
                load_b2(chunks - 1);
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        if (elect_one_sync()) {
            const uint32_t idesc1 = umma_idesc_f16(128, 128);
            const uint32_t idesc2 = umma_idesc_f16(128, n2rows);
            int bs = 0;
            uint32_t bph = 0;
            uint32_t uses1[2] = {0, 0};
            int it = 0;
            auto advance = [&]() {
                if (++bs == nb_gran) {
                    bs = 0;
                    bph ^= 1;
                }
            };
            auto gemm2 = [&](int cc) {
                const int gg = it * chunks + cc;
                const int slot = gg % n_slots;
                mbar_wait(s_aux + oOReady + 8 * slot, static_cast<uint32_t>(gg / n_slots) & 1);
                if (cc == 0) mbar_wait(s_aux + oT2Empty, (it & 1) ^ 1);  // acc2 of the previous tile has been drained
                tc_fence_after();
                for (int h = 0; h < 2; ++h)
                    for (int nh = 0; nh < n2h; ++nh) {
                        mbar_wait(s_aux + oBFull + 8 * bs, bph);
                        tc_fence_after();
                        const uint64_t ad = umma_desc_sw128(sO + slot * kSlot + h * kGran);
                        const uint64_t bd = umma_desc_sw128(sB + bs * kGran);
                        const uint32_t d = tmem_base + 256 + nh * 128;
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_f16(d, ad + 2 * k, bd + 2 * k, idesc2, (cc | h | k) != 0);
                        umma_commit(s_aux + oBEmpty + 8 * bs);
                        advance();
                    }
                umma_commit(s_aux + oOFree + 8 * slot);
            };
            for (int t = blockIdx.x; t < m_tiles; t += gridDim.x, ++it) {
                mbar_wait(s_aux + oAFull, it & 1);
                tc_fence_after();
                for (int c = 0; c < chunks; ++c) {
                    const int st = c & 1;
                    mbar_wait(s_aux + oT1Empty + 8 * st, (uses1[st] & 1) ^ 1);
                    ++uses1[st];
                    tc_fence_after();
                    const uint32_t d = tmem_base + st * 128;
                    for (int kb = 0; kb < kb1; ++kb) {
                        mbar_wait(s_aux + oBFull + 8 * bs, bph);
                        tc_fence_after();
                        const uint64_t ad = umma_desc_sw128(sA + kb * kGran);
                        const uint64_t bd = umma_desc_sw128(sB + bs * kGran);
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_f16(d, ad + 2 * k, bd + 2 * k, idesc1, (kb | k) != 0);
                        umma_commit(s_aux + oBEmpty + 8 * bs);
                        advance();
                    }
                    umma_commit(s_aux + oT1Full + 8 * st);
                    if (c == chunks - 1) umma_commit(s_aux + oAEmpty);
                    if (c >= 1) gemm2(c - 1);
                }
                gemm2(chunks - 1);
                umma_commit(s_aux + oT2Full);
            }
        }
    } else if (warp == 3) {
        // ------------------------------------------------------------ residual producer
        if (lane == 0) {
            int it = 0;
            for (int t = blockIdx.x; t < m_tiles; t += gridDim.x, ++it) {
                const TileXY tc = tail_tile(p, t);
                for (int c = 0; c < chunks; ++c) {
                    const int gg = it * chunks + c;
                    const int slot = gg % n_slots;
                    mbar_wait(s_aux + oOFree + 8 * slot, (static_cast<uint32_t>(gg / n_slots) & 1) ^ 1);
                    const uint32_t full = s_aux + oRFull + 8 * slot;
                    mbar_arrive_expect_tx(full, kSlot);
                    for (int h = 0; h < 2; ++h)
                        tma_load_4d(sO + slot * kSlot + h * kGran, &prob->tmRes, full, c * 128 + h * 64, tc.x0, tc.y0,
                                    tc.n0);
                }
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------ epilogue warpgroups
        const int wg = (warp - 4) >> 2;
        const int wi = warp & 3;  // this warp may touch TMEM lanes [32 * wi, 32 * wi + 32)
        const int et = (threadIdx.x - 128) & 127;
        const int row = wi * 32 + lane;
        const uint32_t bar_id = 1 + wg;
        float2* tab = tables + wg * 128;
        const int tw = p.tw, th = p.th;
        const int rx = row % tw, ry = (row / tw) % th, rn = row / (tw * th);
        uint32_t uses1 = 0;
        int it = 0;
        for (int t = blockIdx.x; t < m_tiles; t += gridDim.x, ++it) {
            const TileXY tc = tail_tile(p, t);
            const int x = tc.x0 + rx, y = tc.y0 + ry, n = tc.n0 + rn;
            const bool valid = x < p.W && y < p.H && n < p.N;
            for (int c = wg; c < chunks; c += 2) {
                const int gg = it * chunks + c;
                const int slot = gg % n_slots;
                named_bar_sync(bar_id, 128);  // everyone is done reading the previous table
                {
                    const int ch = c * 128 + et;
                    tab[et] = make_float2(__ldg(p.scale1 + ch), __ldg(p.shift1 + ch));
                }
                named_bar_sync(bar_id, 128);
                mbar_wait(s_aux + oT1Full + 8 * wg, uses1 & 1);
                ++uses1;
                tc_fence_after();
                mbar_wait(s_aux + oRFull + 8 * slot, static_cast<uint32_t>(gg / n_slots) & 1);
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(wi * 32) << 16) + wg * 128;
#pragma unroll 1
                for (int q4 = 0; q4 < 4; ++q4) {  // quarters of 32 channels
                    const uint32_t buf = sO + slot * kSlot + (q4 >> 1) * kGran;
                    uint32_t v[32];
                    DAFNE_TMEM_LD_X32(taddr + q4 * 32, v);
                    uint4 res[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t src = buf + row * 128 + ((((q4 & 1) * 4 + q) ^ (row & 7)) << 4);
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                     : "=r"(res[q].x), "=r"(res[q].y), "=r"(res[q].z), "=r"(res[q].w)
                                     : "r"(src)
                                     : "memory");
                    }
                    tmem_ld_wait();
                    if (q4 == 3) {
                        // all TMEM reads of this chunk are done: hand the accumulator stage back to the MMA warp
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(s_aux + oT1Empty + 8 * wg);
                    }
                    uint32_t ph[16];
                    tail_math32<true>(v, tab + q4 * 32, res, ph);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t dst = buf + row * 128 + ((((q4 & 1) * 4 + q) ^ (row & 7)) << 4);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(ph[4 * q]),
                                     "r"(ph[4 * q + 1]), "r"(ph[4 * q + 2]), "r"(ph[4 * q + 3])
                                     : "memory");
                    }
                }
                // the staged chunk is read by the async proxy twice: by its TMA store and as the A operand of GEMM 2
                fence_proxy_async_smem();
                named_bar_sync(bar_id, 128);
                if (et == 0) {
                    for (int h = 0; h < 2; ++h)
                        tma_store_4d(&prob->tmOut, sO + slot * kSlot + h * kGran, c * 128 + h * 64, tc.x0, tc.y0, tc.n0);
                    tma_store_commit();
                    mbar_arrive(s_aux + oOReady + 8 * slot);
                    tma_store_wait_read<0>();
                    mbar_arrive(s_aux + oOFree + 8 * slot);
                }
            }
            // ---- acc2 -> mid: this warpgroup's half of the N2 columns, 32 at a time
            mbar_wait(s_aux + oT2Full, it & 1);
            tc_fence_after();
            const int per_wg = N2 >> 1;
            const int subs = per_wg >> 5;
            __half* mrow = p.mid + ((static_cast<size_t>(n) * p.H + y) * p.W + x) * N2;
#pragma unroll 1
            for (int sidx = 0; sidx < subs; ++sidx) {
                const int col0 = wg * per_wg + sidx * 32;
                uint32_t v[32];
                DAFNE_TMEM_LD_X32(tmem_base + (static_cast<uint32_t>(wi * 32) << 16) + 256 + col0, v);
                tmem_ld_wait();
                if (sidx == subs - 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(s_aux + oT2Empty);
                }
                uint32_t ph[16];
#pragma unroll
                for (int c = 0; c < 32; c += 2) {
                    const float a0 = fmaf(__uint_as_float(v[c]), __ldg(p.scale2 + col0 + c), __ldg(p.shift2 + col0 + c));
                    const float a1 =
                        fmaf(__uint_as_float(v[c + 1]), __ldg(p.scale2 + col0 + c + 1), __ldg(p.shift2 + col0 + c + 1));
                    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(ph[c >> 1]) : "f"(a1), "f"(a0));
                }
                if (valid) {
                    uint4* dst = reinterpret_cast<uint4*>(mrow + col0);
#pragma unroll
                    for (int q = 0; q < 4; ++q) dst[q] = make_uint4(ph[4 * q], ph[4 * q + 1], ph[4 * q + 2], ph[4 * q + 3]);
                }
            }
        }
        if (et == 0) tma_store_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------ host side
bool tail_supported(int K1, int N1, int N2) {
    const bool k_ok = K1 == 64 || K1 == 128 || K1 == 256;
    const bool n1_ok = N1 % 256 == 0 && N1 >= 256 && N1 <= 2048;  // an even number of 128-channel chunks
    const bool n2_ok = N2 == 64 || N2 == 128 || N2 == 256;
    return k_ok && n1_ok && n2_ok;
}

static int tail_pow2ceil(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

int tail_plan_build(const TailDesc& d, TailPlan* plan, int num_sms) {
    if (!tail_supported(d.K1, d.N1, d.N2)) {
        set_error("tail_tc: unsupported shape K1=%d N1=%d N2=%d", d.K1, d.N1, d.N2);
        return -1;
    }
    *plan = TailPlan();
    TailParams& p = plan->prob.p;
    p.N = d.N;
    p.H = d.H;
    p.W = d.W;
    p.K1 = d.K1;
    p.N1 = d.N1;
    p.N2 = d.N2;
    p.kb1 = d.K1 / 64;
    p.chunks = d.N1 / 128;
    p.tw = tail_pow2ceil(d.W) < 16 ? tail_pow2ceil(d.W) : 16;
    p.th = tail_pow2ceil(d.H) < 128 / p.tw ? tail_pow2ceil(d.H) : 128 / p.tw;
    p.nb = 128 / (p.tw * p.th);
    p.tiles_x = (d.W + p.tw - 1) / p.tw;
    p.tiles_y = (d.H + p.th - 1) / p.th;
    p.tiles_n = (d.N + p.nb - 1) / p.nb;
    p.m_tiles = p.tiles_x * p.tiles_y * p.tiles_n;
    p.scale1 = d.scale1;
    p.shift1 = d.shift1;
    p.scale2 = d.scale2;
    p.shift2 = d.shift2;
    p.mid = d.mid;
    const uint32_t box_px[4] = {64u, (uint32_t)p.tw, (uint32_t)p.th, (uint32_t)p.nb};
    {
        const uint64_t C = d.K1, W = d.W, H = d.H;
        const uint64_t dims[4] = {C, W, H, (uint64_t)d.N};
        const uint64_t str[3] = {C * 2, W * C * 2, H * W * C * 2};
        if (encode_map(&plan->prob.tmA, d.in, 4, dims, str, box_px, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "tail A")) return -1;
    }
    {
        const uint64_t C = d.N1, W = d.W, H = d.H;
        const uint64_t dims[4] = {C, W, H, (uint64_t)d.N};
        const uint64_t str[3] = {C * 2, W * C * 2, H * W * C * 2};
        if (encode_map(&plan->prob.tmRes, d.residual, 4, dims, str, box_px, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "tail Res"))
            return -1;
        if (encode_map(&plan->prob.tmOut, d.out, 4, dims, str, box_px, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "tail Out"))
            return -1;
    }
    {
        const uint64_t dims[2] = {(uint64_t)d.K1, (uint64_t)d.N1};
        const uint64_t str[1] = {(uint64_t)d.K1 * 2};
        const uint32_t box[2] = {64u, 128u};
        if (encode_map(&plan->prob.tmB1, d.w3, 2, dims, str, box, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "tail B1")) return -1;
    }
    {
        const uint64_t dims[2] = {(uint64_t)d.N1, (uint64_t)d.N2};
        const uint64_t str[1] = {(uint64_t)d.N1 * 2};
        const uint32_t box[2] = {64u, (uint32_t)(d.N2 < 128 ? d.N2 : 128)};
        if (encode_map(&plan->prob.tmB2, d.w1, 2, dims, str, box, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "tail B2")) return -1;
    }
    // shared memory: the activation tile, then as many chunk slots (residual prefetch distance) and weight granules as fit
    const int avail = kTailSmemMax - kAuxBytes - p.kb1 * kGran;
    int slots = 3;
    int gran = (avail - slots * kSlot) / kGran;
    if (gran < 4) {
        slots = 2;
        gran = (avail - slots * kSlot) / kGran;
    }
    if (gran > kMaxGran) {
        gran = kMaxGran;
        if ((avail - gran * kGran) / kSlot >= 4) slots = 4;
    }
    if (gran < 3) {
        set_error("tail_tc: shared memory does not fit K1=%d", d.K1);
        return -1;
    }
    plan->o_slots = slots;
    plan->b_granules = gran;
    plan->smem_bytes = (p.kb1 + gran) * kGran + slots * kSlot + kAuxBytes;
    plan->grid = p.m_tiles < num_sms ? p.m_tiles : num_sms;
    plan->flops = 2.0 * d.N * d.H * d.W * ((double)d.N1 * d.K1 + (double)d.N2 * d.N1);
    return 0;
}

int tail_plan_launch(const TailProblem* dev_prob, const TailPlan& plan, cudaStream_t stream) {
    static DeviceOnce configured;
    int dev;
    if (!configured.get(&dev)) {
        cudaError_t e = cudaFuncSetAttribute(tail_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTailSmemMax);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(tail_tc_kernel): %s", cudaGetErrorString(e));
            return -1;
        }
        configured.set(dev, 1);
    }
    tail_tc_kernel<<<plan.grid, kTailThreads, plan.smem_bytes, stream>>>(dev_prob, plan.b_granules, plan.o_slots);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("tail_tc_kernel launch: %s", cudaGetErrorString(e));
        return -1;
    }
    return 0;
}

}  // namespace dafne
