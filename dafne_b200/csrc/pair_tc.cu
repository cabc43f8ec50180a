// 1x1 / stride-1 convolutions of the deep ResNet stages as a CTA-PAIR tcgen05 GEMM for sm_100a.
//
//   out = act(scale * (in x W^T) + shift (+ residual))        detectron2 v0.5 BottleneckBlock conv1 / conv3 + FrozenBN
//                                                             (+ shortcut) + ReLU via dafne/modeling/backbone/fpn.py:72
//
// Why a second kernel: with one CTA per 128 x 256 tile (conv_tc.cu) the 1x1 convolutions of res4 / res5 re-stream their
// whole weight slab per 128 pixels -- conv1 of res4 (K = 1024) moves 256 KB of activations and 512 KB of weights from
// L2 into shared memory per tile, ~10 TB/s over the launch, which is the L2 -> SM limit, not the tensor pipe. Here two
// CTAs of a cluster (the two SMs of a TPC) work on ONE 256-pixel x 256-channel tile with tcgen05.mma.cta_group::2
// (M = 256): each CTA loads its own 128 pixels of the activations and only HALF of the weight tile (128 of the 256 output
// channels); the tensor cores of both SMs read both halves. Weight traffic per pixel halves, shared-memory fill per SM
// drops by a third, and a stage is 32 KB instead of 48 KB.
//
// Pair protocol (rank 0 = leader):
//   operand stages   both CTAs issue their TMA loads with .cta_group::2 onto the LEADER's full barrier (one
//                    arrive.expect_tx of 2 x 32 KB by the leader's producer); the leader's MMA thread issues the MMAs and
//                    commits with .multicast::cluster to the empty barrier of BOTH CTAs
//   accumulators     2 x 256 TMEM columns in each CTA (rows 128 r ... of the tile in CTA r); the commit of a tile is
//                    multicast to both CTAs' tfull barriers; every epilogue warp of both CTAs arrives on the leader's tempty
//   epilogue         per CTA as in conv_tc.cu / tail_tc.cu: TMEM -> registers -> scale / shift (+ residual from the slot the
//                    TMA put it in) -> ReLU -> fp16 -> swizzled shared memory -> TMA store; slots are owned by ONE
//                    thread (warp 3) that stores a staged chunk, waits for the store to have read it and reloads the
//                    slot with the residual of the chunk that will use it next
// Roles (384 threads per CTA): warp 0 = operand producer, warp 1 = MMA issuer (leader only), warp 2 = TMEM allocator,
// warp 3 = slot manager, warps 4-7 / 8-11 = two epilogue warpgroups (128 output channels of the tile each).
#include <stdio.h>
#include <stdlib.h>

#include "conv_tc.cuh"  // set_error, encode_map, DeviceOnce
#include "pair_tc.cuh"
#include "ptx.cuh"

namespace dafne {

namespace {
constexpr int kGran = 16384;       // 128 rows x 64 channels fp16
constexpr int kStage = 2 * kGran;  // A (own 128 pixels) + B (own 128 output channels)
constexpr int kSlot = 2 * kGran;   // one output / residual chunk: 128 pixels x 128 channels
constexpr int kPairThreads = 384;
constexpr int kMaxStages = 8, kMaxSlots = 4;
constexpr int kAuxBars = 512;
constexpr int kAuxBytes = kAuxBars + 2 * 128 * 8;
constexpr int kPairSmemMax = 232448;  // 227 KB
constexpr int oFull = 0, oEmpty = 64, oTFull = 128, oTEmpty = 144, oRFull = 160, oOReady = 192, oOFree = 224,
              oTmemPtr = 256;

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load of a CTA pair: the data lands in THIS CTA's shared memory, the bytes are counted on `bar_cluster`, which
// may be the peer's barrier
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
        "[%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(src), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[smem of both CTAs: 2 x 128 rows] * B[smem of both CTAs: 2 x N/2 rows]^T
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs once all MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
    const uint16_t mask = 3;
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
        "h"(mask)
        : "memory");
}

// act(acc * scale + shift (+ residual)) for 32 channels -> 16 packed fp16 pairs
template <bool HAS_RES>
__device__ __forceinline__ void pair_math32(const uint32_t (&v)[32], const float2* tab, const uint4 (&res)[4], bool relu,
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
        if (relu)
            asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(a1), "f"(a0));
        else
            asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(a1), "f"(a0));
        packed[c >> 1] = h;
    }
}
}  // namespace

// pair tile index -> (pixel-pair tile, channel tile): the channel tile is the fast index, so pairs that run side by side
// share the activation rows through L2
__device__ __forceinline__ void pair_tile(const PairParams& p, int pt, int* mt, int* nt) {
    *nt = pt % p.n_tiles;
    const int m = pt / p.n_tiles;
    *mt = p.reverse_m ? p.m_pairs - 1 - m : m;
}

template <bool HAS_RES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1)
    pair_tc_kernel(const PairProblem* __restrict__ prob, int stages, int n_slots) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const PairParams& p = prob->p;
    const int kbs = p.kbs, n_tiles = p.n_tiles, total = p.total;
    (void)n_tiles;

    const uint32_t smem_base = smem_u32(smem);
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    // carve-up: [operand stages: stages x (A | B)] [chunk slots: n_slots x 32 KB] [aux]
    const uint32_t sT = smem_base;
    const uint32_t sO = sT + stages * kStage;
    const uint32_t s_aux = sO + n_slots * kSlot;
    uint8_t* aux = smem + stages * kStage + n_slots * kSlot;
    volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(aux + oTmemPtr);
    float2* tables = reinterpret_cast<float2*>(aux + kAuxBars);

    if (threadIdx.x == 0 && (smem_base & 1023u) != 0) {
        printf("dafne pair_tc: dynamic smem base not 1024-aligned (%u)\n", smem_base);
        __trap();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&prob->tmA);
        tma_prefetch_desc(&prob->tmB);
        tma_prefetch_desc(&prob->tmOut);
        if (HAS_RES) tma_prefetch_desc(&prob->tmRes);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kMaxStages; ++i) {
            mbar_init(s_aux + oFull + 8 * i, 1);   // used in the leader only
            mbar_init(s_aux + oEmpty + 8 * i, 1);  // one multicast commit of the leader's MMA thread
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(s_aux + oTFull + 8 * i, 1);
            mbar_init(s_aux + oTEmpty + 8 * i, 16);  // used in the leader only: 8 epilogue warps of each CTA
        }
        for (int i = 0; i < kMaxSlots; ++i) {
            mbar_init(s_aux + oRFull + 8 * i, 1);
            mbar_init(s_aux + oOReady + 8 * i, 1);
            mbar_init(s_aux + oOFree + 8 * i, 1);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc_pair(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), 512);
        tmem_relinquish_pair();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // the peer's barriers exist before anything of this CTA signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ------------------------------------------------------------ operand producer (both CTAs)
        if (elect_one_sync()) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t full0 = mapa_rank(s_aux + oFull, 0);  // the leader's full barriers
            for (int pt = pair; pt < total; pt += npairs) {
                int nt, mt;
                pair_tile(p, pt, &mt, &nt);
                const int row0 = mt * 256 + static_cast<int>(rank) * 128;
                const int wrow0 = nt * 256 + static_cast<int>(rank) * 128;
                for (int kb = 0; kb < kbs; ++kb) {
                    mbar_wait(s_aux + oEmpty + 8 * stage, phase ^ 1);
                    if (rank == 0) mbar_arrive_expect_tx(s_aux + oFull + 8 * stage, 2 * kStage);
                    const uint32_t dst = sT + stage * kStage;
                    tma_load_2d_pair(dst, &prob->tmA, full0 + 8 * stage, kb * 64, row0);
                    tma_load_2d_pair(dst + kGran, &prob->tmB, full0 + 8 * stage, kb * 64, wrow0);
                    if (++stage == stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (leader CTA only)
        if (rank == 0 && elect_one_sync()) {
            constexpr uint32_t idesc = umma_idesc_f16(256, 256);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int pt = pair; pt < total; pt += npairs, ++it) {
                const int acc = it & 1;
                mbar_wait(s_aux + oTEmpty + 8 * acc, ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + acc * 256;
                for (int kb = 0; kb < kbs; ++kb) {
                    mbar_wait(s_aux + oFull + 8 * stage, phase);
                    tc_fence_after();
                    const uint64_t ad = umma_desc_sw128(sT + stage * kStage);
                    const uint64_t bd = umma_desc_sw128(sT + stage * kStage + kGran);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16_pair(d, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
                    umma_commit_pair(s_aux + oEmpty + 8 * stage);
                    if (++stage == stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit_pair(s_aux + oTFull + 8 * acc);
            }
        }
    } else if (warp == 3) {
        // ------------------------------------------------------------ slot manager: stores, then reloads / frees the slot
        if (lane == 0) {
            const int my_tiles = pair < total ? (total - pair + npairs - 1) / npairs : 0;
            const int nchunks = 2 * my_tiles;
            auto chunk_coords = [&](int gg, int* col, int* row) {
                int nt, mt;
                pair_tile(p, pair + (gg >> 1) * npairs, &mt, &nt);
                *col = nt * 256 + (gg & 1) * 128;
                *row = mt * 256 + static_cast<int>(rank) * 128;
            };
            auto load_res = [&](int gg) {
                const int slot = gg % n_slots;
                int col, row;
                chunk_coords(gg, &col, &row);
                const uint32_t full = s_aux + oRFull + 8 * slot;
                mbar_arrive_expect_tx(full, kSlot);
                tma_load_2d(sO + slot * kSlot, &prob->tmRes, full, col, row);
                tma_load_2d(sO + slot * kSlot + kGran, &prob->tmRes, full, col + 64, row);
            };
            if (HAS_RES)
                for (int gg = 0; gg < n_slots && gg < nchunks; ++gg) load_res(gg);
            for (int gg = 0; gg < nchunks; ++gg) {
                const int slot = gg % n_slots;
                int col, row;
                chunk_coords(gg, &col, &row);
                mbar_wait(s_aux + oOReady + 8 * slot, static_cast<uint32_t>(gg / n_slots) & 1);
                tma_store_2d(&prob->tmOut, sO + slot * kSlot, col, row);
                tma_store_2d(&prob->tmOut, sO + slot * kSlot + kGran, col + 64, row);
                tma_store_commit();
                tma_store_wait_read<0>();
                if (HAS_RES) {
                    if (gg + n_slots < nchunks) load_res(gg + n_slots);
                } else {
                    mbar_arrive(s_aux + oOFree + 8 * slot);
                }
            }
            tma_store_wait_all();
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------ epilogue warpgroups
        const int wg = (warp - 4) >> 2;
        const int wi = warp & 3;  // this warp may touch TMEM lanes [32 * wi, 32 * wi + 32)
        const int et = (threadIdx.x - 128) & 127;
        const int row = wi * 32 + lane;
        const uint32_t bar_id = 1 + wg;
        float2* tab = tables + wg * 128;
        const bool relu = p.relu != 0;
        const uint32_t tempty0 = mapa_rank(s_aux + oTEmpty, 0);  // the leader's tempty barriers
        int it = 0;
        for (int pt = pair; pt < total; pt += npairs, ++it) {
            const int nt = pt % n_tiles;
            const int acc = it & 1;
            const int gg = it * 2 + wg;
            const int slot = gg % n_slots;
            named_bar_sync(bar_id, 128);  // everyone is done reading the previous table
            {
                const int ch = nt * 256 + wg * 128 + et;
                tab[et] = make_float2(__ldg(p.scale + ch), __ldg(p.shift + ch));
            }
            named_bar_sync(bar_id, 128);
            mbar_wait(s_aux + oTFull + 8 * acc, (it >> 1) & 1);
            tc_fence_after();
            if (HAS_RES)
                mbar_wait(s_aux + oRFull + 8 * slot, static_cast<uint32_t>(gg / n_slots) & 1);
            else
                mbar_wait(s_aux + oOFree + 8 * slot, (static_cast<uint32_t>(gg / n_slots) & 1) ^ 1);
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(wi * 32) << 16) + acc * 256 + wg * 128;
#pragma unroll 1
            for (int q4 = 0; q4 < 4; ++q4) {  // quarters of 32 channels
                const uint32_t buf = sO + slot * kSlot + (q4 >> 1) * kGran;
                uint32_t v[32];
                DAFNE_TMEM_LD_X32(taddr + q4 * 32, v);
                uint4 res[4];
                if (HAS_RES) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t src = buf + row * 128 + ((((q4 & 1) * 4 + q) ^ (row & 7)) << 4);
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                     : "=r"(res[q].x), "=r"(res[q].y), "=r"(res[q].z), "=r"(res[q].w)
                                     : "r"(src)
                                     : "memory");
                    }
                }
                tmem_ld_wait();
                if (q4 == 3) {
                    // all TMEM reads of this tile are done: hand the accumulator stage back to the leader's MMA thread
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(tempty0 + 8 * acc);
                }
                uint32_t ph[16];
                pair_math32<HAS_RES>(v, tab + q4 * 32, res, relu, ph);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t dst = buf + row * 128 + ((((q4 & 1) * 4 + q) ^ (row & 7)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(ph[4 * q]), "r"(ph[4 * q + 1]),
                                 "r"(ph[4 * q + 2]), "r"(ph[4 * q + 3])
                                 : "memory");
                }
            }
            fence_proxy_async_smem();  // the staged chunk is read by the TMA store (async proxy)
            named_bar_sync(bar_id, 128);
            if (et == 0) mbar_arrive(s_aux + oOReady + 8 * slot);
        }
    }

    // nobody leaves (or frees TMEM) while the peer may still signal this CTA's barriers or read its operands
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------ host side
bool pair_supported(int K, int N) { return K >= 64 && K % 64 == 0 && N >= 256 && N % 256 == 0; }

int pair_plan_build(const PairDesc& d, PairPlan* plan, int num_sms) {
    if (!pair_supported(d.K, d.N) || d.M <= 0 || d.M > 0x7fffff00LL || !d.in || !d.w || !d.out || !d.scale || !d.shift) {
        set_error("pair_tc: unsupported problem M=%lld K=%d N=%d", d.M, d.K, d.N);
        return -1;
    }
    *plan = PairPlan();
    PairParams& p = plan->prob.p;
    p.K = d.K;
    p.N = d.N;
    p.kbs = d.K / 64;
    p.n_tiles = d.N / 256;
    p.m_pairs = static_cast<int>((d.M + 255) / 256);
    p.total = p.m_pairs * p.n_tiles;
    p.relu = d.relu;
    p.reverse_m = d.reverse_m;
    p.scale = d.scale;
    p.shift = d.shift;
    const uint32_t box[2] = {64u, 128u};
    {
        const uint64_t dims[2] = {(uint64_t)d.K, (uint64_t)d.M};
        const uint64_t str[1] = {(uint64_t)d.K * 2};
        if (encode_map(&plan->prob.tmA, d.in, 2, dims, str, box, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "pair A")) return -1;
    }
    {
        const uint64_t dims[2] = {(uint64_t)d.K, (uint64_t)d.N};
        const uint64_t str[1] = {(uint64_t)d.K * 2};
        if (encode_map(&plan->prob.tmB, d.w, 2, dims, str, box, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "pair B")) return -1;
    }
    {
        const uint64_t dims[2] = {(uint64_t)d.N, (uint64_t)d.M};
        const uint64_t str[1] = {(uint64_t)d.N * 2};
        if (encode_map(&plan->prob.tmOut, d.out, 2, dims, str, box, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "pair Out")) return -1;
        if (d.residual &&
            encode_map(&plan->prob.tmRes, d.residual, 2, dims, str, box, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "pair Res"))
            return -1;
    }
    plan->has_res = d.residual != nullptr;
    // shared memory: residual convolutions keep three chunk slots (residual prefetch distance), the others two
    plan->slots = plan->has_res ? 3 : 2;
    if (const char* ev = getenv("DAFNE_PAIR_SLOTS")) plan->slots = atoi(ev);
    if (plan->slots < 2) plan->slots = 2;
    if (plan->slots > kMaxSlots) plan->slots = kMaxSlots;
    plan->stages = (kPairSmemMax - kAuxBytes - plan->slots * kSlot) / kStage;
    if (plan->stages > kMaxStages) plan->stages = kMaxStages;
    if (plan->stages < 2) {
        set_error("pair_tc: shared memory does not fit");
        return -1;
    }
    plan->smem_bytes = plan->stages * kStage + plan->slots * kSlot + kAuxBytes;
    const int pairs = num_sms / 2;
    plan->grid = 2 * (p.total < pairs ? p.total : pairs);
    plan->flops = 2.0 * static_cast<double>(d.M) * d.N * d.K;
    plan->bytes = 2.0 * (static_cast<double>(d.M) * d.K + static_cast<double>(d.N) * d.K +
                         static_cast<double>(d.M) * d.N * (plan->has_res ? 2 : 1));
    return 0;
}

int pair_plan_launch(const PairProblem* dev_prob, const PairPlan& plan, cudaStream_t stream) {
    static DeviceOnce configured;
    int dev;
    if (!configured.get(&dev)) {
        cudaError_t e = cudaFuncSetAttribute(pair_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmemMax);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(pair_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmemMax);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(pair_tc_kernel): %s", cudaGetErrorString(e));
            return -1;
        }
        configured.set(dev, 1);
    }
    if (plan.has_res)
        pair_tc_kernel<true><<<plan.grid, kPairThreads, plan.smem_bytes, stream>>>(dev_prob, plan.stages, plan.slots);
    else
        pair_tc_kernel<false><<<plan.grid, kPairThreads, plan.smem_bytes, stream>>>(dev_prob, plan.stages, plan.slots);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("pair_tc_kernel launch: %s", cudaGetErrorString(e));
        return -1;
    }
    return 0;
}

}  // namespace dafne
