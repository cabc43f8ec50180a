// im2col-free NHWC convolution (1x1 / 3x3, stride 1 / 2) as a persistent, warp-specialised tcgen05 kernel for sm_100a.
//
// Replaces, for the inference hot path, every cuDNN conv2d the reference reaches through torch:
//   ResNet bottlenecks + FrozenBN (+ReLU, +shortcut)   detectron2 v0.5 BottleneckBlock via dafne/modeling/backbone/fpn.py:72
//   FPN lateral / output convs, P6 / P7                 dafne/modeling/backbone/fpn.py:16-37,83-90
//   head tower convs and prediction convs               dafne/modeling/dafne/dafne.py:287-348, 209-230, 388-414, 462-471
//
// GEMM view: M = output pixels (tile = 128 pixels = one tw x th x nb patch), N = Cout (tile BLOCK_N), K = taps x Cin in
// blocks of 64 channels (one 128-byte swizzle row). For each (tap, channel block) ONE 4-D TMA box load of the shifted
// input patch lands directly in the canonical K-major SWIZZLE_128B operand layout; out-of-image coordinates are
// zero-filled by TMA, which is exactly the conv zero padding. Nothing im2col-shaped ever exists in HBM.
//
// Halo box (3x3 stride 1 with a narrow N tile, where the kernel is bound by L2 -> shared-memory fills, not by the
// tensor pipe: the prediction convs, res2's 3x3): the pixel tile is 8 wide x 16 high and ONE box of 18 rows x 10
// pixels (the tile plus a one-pixel border) per channel block serves all nine taps: the operand of tap (dy, dx) is
// that box read from pixel (1 + dy) * 10 + (1 + dx) on, eight consecutive pixels per 8-row group, groups 10 pixels
// (1280 B) apart. Such a descriptor start is not on a 1024-byte atom boundary; it works because both TMA and the
// tensor core derive the 128-byte swizzle from the shared-memory ADDRESS bits [7, 10) (measured: results bit-equal to
// the atom-aligned forms with base_offset = 0, wrong with base_offset = start >> 7). One A load per channel block
// instead of nine. (The "row-shared" predecessor -- 8-pixel-wide boxes, one per dx, starts moving by whole atoms --
// is kept for A/B runs, DAFNE_CONV_TAPS=rows.)
//
// Roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one thread), warp 2 = TMEM allocator,
// warp 3 = residual producer, warps 4-7 = epilogue (TMEM -> registers -> scale/shift/residual/ReLU -> fp16 -> swizzled
// smem -> TMA store). Two TMEM accumulator stages let the epilogue of tile i overlap the MMAs of tile i+1.
//
// Residual (the bottleneck shortcut, detectron2 BottleneckBlock `out += shortcut`): the 128 px x 64 ch residual chunk
// is TMA-loaded by warp 3 into the very ring slot the epilogue later stages its output in (same box, same swizzle), so
// the epilogue adds it from shared memory in place and the slot goes load -> add/ReLU -> store -> free. The loads run
// a whole ring ahead of the epilogue; nothing on that path is an uncoalesced per-thread global load.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "conv_tc.cuh"
#include "ptx.cuh"

namespace dafne {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }

// EPI_WGS = number of epilogue warpgroups. 1: compute-bound shapes (deep K). 2: HBM-bound shapes (K <= 256, the
// epilogue is the critical path): two warpgroups drain alternate tiles (one TMEM accumulator stage each) so two tile
// epilogues are in flight. The number of operand stages and of 16 KB epilogue ring slots per warpgroup is chosen at
// launch (conv_smem_config) -- shared memory is carved up at run time.
template <int BLOCK_N, int EPI_WGS>
struct ConvCfg {
    static constexpr int A_BYTES = 128 * 128;      // 128 pixels x 64 ch fp16
    static constexpr int B_BYTES = BLOCK_N * 128;  // BLOCK_N couts x 64 ch fp16
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int SLOT_BYTES = 16384;  // 128 pixels x 64 ch fp16, one epilogue chunk
    static constexpr int AUX_HDR = 512;       // barriers (2 x MAX_RING residual slots at +256 / +352), tmem pointer, tile offsets
    // + (scale, shift) pairs per warpgroup; MODE 2 (tower convs: bias only) keeps the shifts alone
    static constexpr int aux_bytes(int mode) { return AUX_HDR + EPI_WGS * BLOCK_N * (mode == 2 ? 4 : 8); }
    static constexpr int AUX_BYTES = AUX_HDR + EPI_WGS * BLOCK_N * 8;
    static constexpr int TMEM_COLS = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
    static constexpr int THREADS = 128 + 128 * EPI_WGS;
    static constexpr int MAX_STAGES = 8, MAX_RING = 6;
    // row-shared taps: A box = 18 rows x 8 px x 64 ch (18 KB, padded to 19 KB so B stays 1024-aligned) + B of 3 taps
    static constexpr int A_RS_BOX_BYTES = 18 * 1024;
    static constexpr int A_RS_BYTES = 19 * 1024;
    static constexpr int STAGE_BYTES_RS = A_RS_BYTES + 3 * B_BYTES;
    // halo box: one 18-row x 10-pixel box per channel block serves all nine taps through descriptor starts that are
    // NOT atom-aligned (stride between 8-row groups = 10 pixels = 1280 B)
    static constexpr int A_HALO_BOX_BYTES = 18 * 10 * 128;
    static constexpr int A_HALO_BYTES = 23 * 1024;
    static constexpr int STAGE_BYTES_HALO = A_HALO_BYTES + 9 * B_BYTES;
    // mode 3 = halo box with the whole weight tensor of the problem RESIDENT in shared memory (all taps x channel
    // blocks, <= kMaxResidentB bytes): it is loaded once per run of tiles that share it, a stage is the A box alone
    // mode 5 = halo boxes in their OWN two-slot ring in front of the stages (A5_RING_BYTES + the GroupNorm table), a
    // stage is the weight tile of ONE tap: the A box of a channel block is loaded once and serves nine stages
    static constexpr int A5_SLOTS = 3;
    static constexpr int A5_RING_BYTES = A5_SLOTS * A_HALO_BYTES;  // 69 KB
    static constexpr int GN_TAB_BYTES = 256 * 8;                    // [2][32] (mean, rstd) per group, rounded up to 2 KB
    static constexpr int A5_REGION_BYTES = A5_RING_BYTES + GN_TAB_BYTES;  // 71 KB, a multiple of 1024
    __host__ __device__ static constexpr int stage_bytes(int row_shared) {
        return row_shared == 5 ? B_BYTES
                               : (row_shared == 3 ? A_HALO_BYTES
                                                  : (row_shared == 2 ? STAGE_BYTES_HALO
                                                                     : (row_shared ? STAGE_BYTES_RS : STAGE_BYTES)));
    }
    static constexpr int smem_bytes(int stages, int ring, int row_shared = 0, int breg_bytes = 0, int mode = 0) {
        return breg_bytes + stages * stage_bytes(row_shared) +
               (BLOCK_N >= 64 ? EPI_WGS * ring * SLOT_BYTES : 0) + aux_bytes(mode);
    }
};
constexpr int kMaxSmem = 232448;  // 227 KB
constexpr int kMaxResidentB = 72 * 1024;

// Sum 16 per-lane values across the warp; lane l returns the total of value index
// 8*b4 + 4*b3 + 2*b2 + b1 (b_i = bit i of l). 16 shuffles instead of 80.
template <typename T>
__device__ __forceinline__ T warp_reduce16_scatter(const T (&v)[16], uint32_t lane) {
    const uint32_t full = 0xffffffffu;
    T a[8], b[4], c[2];
    bool hi = lane & 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        T send = hi ? v[i] : v[i + 8];
        T keep = hi ? v[i + 8] : v[i];
        a[i] = keep + __shfl_xor_sync(full, send, 16);
    }
    hi = lane & 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        T send = hi ? a[i] : a[i + 4];
        T keep = hi ? a[i + 4] : a[i];
        b[i] = keep + __shfl_xor_sync(full, send, 8);
    }
    hi = lane & 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        T send = hi ? b[i] : b[i + 2];
        T keep = hi ? b[i + 2] : b[i];
        c[i] = keep + __shfl_xor_sync(full, send, 4);
    }
    hi = lane & 2;
    T send = hi ? c[0] : c[1];
    T keep = hi ? c[1] : c[0];
    T d = keep + __shfl_xor_sync(full, send, 2);
    d += __shfl_xor_sync(full, d, 1);
    return d;
}

// tile index inside a problem -> (m tile, n tile): the n tile is the FAST index, so CTAs that run side by side share
// the activation tile (second reader hits L2) instead of re-reading the whole input once per n tile.
struct TileCoord {
    int x0, y0, n0, nt;
};
__device__ __forceinline__ TileCoord tile_coord(const ConvParams& p, int lt) {
    TileCoord c;
    c.nt = lt % p.n_tiles;
    int mt = lt / p.n_tiles;
    if (p.reverse_m) mt = p.m_tiles - 1 - mt;
    const int tx = mt % p.tiles_x;
    const int r = mt / p.tiles_x;
    c.x0 = tx * p.tw;
    c.y0 = (r % p.tiles_y) * p.th;
    c.n0 = (r / p.tiles_y) * p.nb;
    return c;
}

// out[c], out[c+1] = fp16(relu?(acc * scale + shift (+ residual))) for 32 channels (HALF a chunk: with a whole chunk
// in flight -- 64 accumulators + 32 residual + 32 packed registers -- the epilogue needed 196 registers; in halves it
// needs 145, which fits the 168 a 384-thread CTA starts with). The ReLU rides on
// the conversion (cvt.rn.relu.f16x2.f32), the (scale, shift) pairs come as one 128-bit shared-memory broadcast per two
// channels. Compile-time RELU / HAS_RES: the epilogue of the HBM-bound convolutions is issue-bound, so nothing that is
// switched off may cost an instruction.
template <bool RELU, bool HAS_RES>
__device__ __forceinline__ void epilogue_half_math(const uint32_t (&v)[32], const float2* tab, const uint4 (&res)[4],
                                                   uint32_t (&packed)[16]) {
    const uint32_t* rw = reinterpret_cast<const uint32_t*>(res);
#pragma unroll
    for (int c = 0; c < 32; c += 2) {
        const float4 tb = *reinterpret_cast<const float4*>(&tab[c]);  // (scale, shift) x 2
        float a0 = fmaf(__uint_as_float(v[c]), tb.x, tb.y);
        float a1 = fmaf(__uint_as_float(v[c + 1]), tb.z, tb.w);
        if (HAS_RES) {
            const float2 rf = __half22float2(*reinterpret_cast<const __half2*>(&rw[c >> 1]));
            a0 += rf.x;
            a1 += rf.y;
        }
        uint32_t h;
        if (RELU)
            asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(a1), "f"(a0));
        else
            asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(a1), "f"(a0));
        packed[c >> 1] = h;
    }
}

// Bias-only variant for the tower convolutions (MODE 2: no scale, no residual, no ReLU -- GroupNorm follows): the table
// holds the 64 shifts of the chunk.
__device__ __forceinline__ void epilogue_chunk_bias(const uint32_t (&v)[64], const float* tab, uint32_t (&packed)[32]) {
#pragma unroll
    for (int c = 0; c < 64; c += 4) {
        const float4 sh = *reinterpret_cast<const float4*>(&tab[c]);
        const float a0 = __uint_as_float(v[c]) + sh.x, a1 = __uint_as_float(v[c + 1]) + sh.y;
        const float a2 = __uint_as_float(v[c + 2]) + sh.z, a3 = __uint_as_float(v[c + 3]) + sh.w;
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(packed[c >> 1]) : "f"(a1), "f"(a0));
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(packed[(c >> 1) + 1]) : "f"(a3), "f"(a2));
    }
}

// MODE: what the epilogue does besides scale/shift/ReLU -- 0 nothing, 1 residual add (TMA ring when res_tma, else
// per-thread loads of a nearest-upsampled residual), 2 GroupNorm statistics of the fp16-rounded output.
template <int BLOCK_N, int EPI_WGS, int MODE>
__global__ void __launch_bounds__(128 + 128 * EPI_WGS, 1)
    conv_tc_kernel(const ConvProblem* __restrict__ probs, int nprob, int total_tiles, int stages, int ring,
                   int res_tma_arg, int row_shared, int breg_bytes) {
    const int res_tma = MODE == 1 ? res_tma_arg : 0;
    using Cfg = ConvCfg<BLOCK_N, EPI_WGS>;
    // mode 5 (halo-box ring + GroupNorm on load) exists only in the variant the tower convolutions use: compiled into
    // every variant, its transform role made ptxas spill in the 88-register producer side of the two-warpgroup kernels
    constexpr bool kHalo5 = BLOCK_N == 256 && EPI_WGS == 1 && MODE == 2;
    const bool halo5 = kHalo5 && row_shared == 5;
    extern __shared__ __align__(1024) uint8_t smem[];

    const uint32_t smem_base = smem_u32(smem);
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (threadIdx.x == 0 && (smem_base & 1023u) != 0) {
        printf("dafne conv_tc: dynamic smem base not 1024-aligned (%u)\n", smem_base);
        __trap();
    }

    // carve-up: [resident weights (mode 3)] [stages x (A | B)] [EPI_WGS x ring x 16 KB epilogue slots] [512 B header]
    // [(scale, shift) tables]
    const int epi_bytes = BLOCK_N >= 64 ? EPI_WGS * ring * Cfg::SLOT_BYTES : 0;
    const uint32_t s_bres = smem_base;
    const uint32_t s_tiles = smem_base + breg_bytes;
    const int stage_bytes = Cfg::stage_bytes(row_shared);
    const uint32_t s_epi = s_tiles + stages * stage_bytes;
    const uint32_t s_aux = s_epi + epi_bytes;
    uint8_t* aux = smem + breg_bytes + stages * stage_bytes + epi_bytes;
    const uint32_t bar_full = s_aux;             // 8 x 8 B
    const uint32_t bar_empty = s_aux + 64;       // 8 x 8 B
    const uint32_t bar_tfull = s_aux + 128;      // 2 x 8 B
    const uint32_t bar_tempty = s_aux + 144;     // 2 x 8 B
    volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(aux + 160);
    int* s_begin = reinterpret_cast<int*>(aux + 168);  // nprob + 1 tile offsets (<= 17 ints)
    constexpr int RS = Cfg::MAX_RING;            // barrier slots per warpgroup
    const uint32_t bar_rfull = s_aux + 256;      // [2][MAX_RING] x 8 B: residual chunk landed in the slot
    const uint32_t bar_rempty = s_aux + 352;     // [2][MAX_RING] x 8 B: the slot's output store has been read out
    const uint32_t bar_bfull = s_aux + 448;      // resident weights landed
    // mode 5 (never together with the residual ring: MODE 2 has none), 3 x 8 B each, in the residual barriers' space
    const uint32_t bar_afull = s_aux + 256;      // halo box landed
    const uint32_t bar_aready = s_aux + 280;     // ... and GroupNorm + ReLU applied to it
    const uint32_t bar_aempty = s_aux + 304;     // its nine taps have been multiplied
    float2* s_tab_all = reinterpret_cast<float2*>(aux + Cfg::AUX_HDR);

    if (warp == 0 && lane < nprob) {
        tma_prefetch_desc(&probs[lane].tmA[0]);
        tma_prefetch_desc(&probs[lane].tmB);
        if (BLOCK_N >= 64) tma_prefetch_desc(&probs[lane].tmOut);
        if (BLOCK_N >= 64 && res_tma) tma_prefetch_desc(&probs[lane].tmRes);
    }
    if (warp == 3 && lane <= nprob) s_begin[lane] = lane < nprob ? probs[lane].p.tile_begin : total_tiles;
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < stages; ++i) {
            mbar_init(bar_full + 8 * i, 1);
            mbar_init(bar_empty + 8 * i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_tfull + 8 * i, 1);
            mbar_init(bar_tempty + 8 * i, 4);  // one arrive per epilogue warp
        }
        if (!halo5)
            for (int i = 0; i < 2 * RS; ++i) {
                mbar_init(bar_rfull + 8 * i, 1);
                mbar_init(bar_rempty + 8 * i, 1);
            }
        mbar_init(bar_bfull, 1);
        if (halo5)
            for (int i = 0; i < Cfg::A5_SLOTS; ++i) {
                mbar_init(bar_afull + 8 * i, 1);
                mbar_init(bar_aready + 8 * i, 1);
                mbar_init(bar_aempty + 8 * i, 1);
            }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    // EPI_WGS == 2: 384 threads start with 168 registers each; the producer / MMA warpgroup hands part of its share to
    // the two epilogue warpgroups: 128 x 88 + 256 x 208 = 384 x 168 EXACTLY -- setmaxnreg.inc blocks until the CTA's own
    // pool holds the registers, so asking for more than the other side released hangs the kernel (88 / 208 and the
    // earlier 56 / 224 balance; 96 / 208 does not). ptxas allocates each side under its setmaxnreg
    // value: with the earlier 56 / 224 split the PRODUCER warps spilled (84 B stores, 412 B loads per thread, inside
    // the per-tile loop that feeds the HBM-bound 1x1 convolutions).
    if (warp < 4) {
        if constexpr (EPI_WGS == 2) asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    }
    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (operands)
        if (DAFNE_ONE_THREAD(2, lane)) {
            int stage = 0;
            uint32_t phase = 0;
            int g = 0;
            long long bkey = -1;  // mode 3: which weights are resident
            int last_stage = -1;
            uint32_t last_phase = 0;
            if (halo5) {
                // Halo boxes run two channel blocks ahead of the weight stages (three slots): the box of block cb + 2 (of
                // the next tile after the last block) is requested right after the nine weight tiles of block cb -- the
                // slot it takes, block cb - 1's, is free by then, so the request never stalls the weight stream -- and
                // has about thirteen stages' worth of tensor work (6 600 cycles) to arrive from HBM and be transformed.
                int as = 0;  // A ring position of the next box to request
                uint32_t aph = 0;
                int at = blockIdx.x, acb = 0, ag = 0;  // (tile, channel block, problem) of the next box to request
                auto request_box = [&]() {
                    if (at >= total_tiles) return;
                    while (at >= s_begin[ag + 1]) ++ag;
                    const ConvProblem* pa = probs + ag;
                    const TileCoord ta = tile_coord(pa->p, at - s_begin[ag]);
                    mbar_wait(bar_aempty + 8 * as, aph ^ 1);
                    mbar_arrive_expect_tx(bar_afull + 8 * as, Cfg::A_HALO_BOX_BYTES);
                    tma_load_4d(s_bres + as * Cfg::A_HALO_BYTES, &pa->tmA[0], bar_afull + 8 * as, acb * 64, ta.x0 - 1,
                                ta.y0 - 1, ta.n0);
                    if (++as == Cfg::A5_SLOTS) {
                        as = 0;
                        aph ^= 1;
                    }
                    if (++acb == pa->p.cin_blocks) {
                        acb = 0;
                        at += gridDim.x;
                    }
                };
                request_box();
                request_box();
                for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                    while (t >= s_begin[g + 1]) ++g;
                    const ConvProblem* pr = probs + g;
                    const ConvParams& p = pr->p;
                    const TileCoord tc = tile_coord(p, t - s_begin[g]);
                    for (int cb = 0; cb < p.cin_blocks; ++cb) {
#pragma unroll 1
                        for (int tap = 0; tap < 9; ++tap) {
                            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                            const uint32_t full = bar_full + 8 * stage;
                            mbar_arrive_expect_tx(full, Cfg::B_BYTES);
                            tma_load_2d(s_tiles + stage * stage_bytes, &pr->tmB, full, tap * p.Cin + cb * 64,
                                        tc.nt * BLOCK_N);
                            if (++stage == stages) {
                                stage = 0;
                                phase ^= 1;
                            }
                        }
                        // AFTER this block's weight tiles: its slot's predecessor (block cb - 1) is done by now, so
                        // this never blocks the weight stream behind the tensor core
                        request_box();
                    }
                }
            } else
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                while (t >= s_begin[g + 1]) ++g;
                const ConvProblem* pr = probs + g;
                const ConvParams& p = pr->p;
                const TileCoord tc = tile_coord(p, t - s_begin[g]);
                const int num_taps = p.num_taps, cin_blocks = p.cin_blocks, Cin = p.Cin;
                if (row_shared == 3) {
                    // weights of this (problem, n tile): resident, reloaded only when they change -- after the MMAs
                    // that still read the old ones have drained (the last stage filled has been consumed)
                    const long long key = (static_cast<long long>(reinterpret_cast<uintptr_t>(p.w_id)) << 8) | tc.nt;
                    if (key != bkey) {
                        if (last_stage >= 0) mbar_wait(bar_empty + 8 * last_stage, last_phase);
                        mbar_arrive_expect_tx(bar_bfull, 9 * cin_blocks * Cfg::B_BYTES);
                        for (int cb = 0; cb < cin_blocks; ++cb)
                            for (int tap = 0; tap < 9; ++tap)
                                tma_load_2d(s_bres + (cb * 9 + tap) * Cfg::B_BYTES, &pr->tmB, bar_bfull,
                                            tap * Cin + cb * 64, tc.nt * BLOCK_N);
                        bkey = key;
                    }
                    for (int cb = 0; cb < cin_blocks; ++cb) {
                        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                        const uint32_t full = bar_full + 8 * stage;
                        mbar_arrive_expect_tx(full, Cfg::A_HALO_BOX_BYTES);
                        tma_load_4d(s_tiles + stage * stage_bytes, &pr->tmA[0], full, cb * 64, tc.x0 - 1, tc.y0 - 1,
                                    tc.n0);
                        last_stage = stage;
                        last_phase = phase;
                        if (++stage == stages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    continue;
                }
                if (row_shared == 2) {
                    for (int cb = 0; cb < cin_blocks; ++cb) {
                        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                        const uint32_t full = bar_full + 8 * stage;
                        mbar_arrive_expect_tx(full, Cfg::A_HALO_BOX_BYTES + 9 * Cfg::B_BYTES);
                        const uint32_t sA = s_tiles + stage * stage_bytes;
                        tma_load_4d(sA, &pr->tmA[0], full, cb * 64, tc.x0 - 1, tc.y0 - 1, tc.n0);
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap)
                            tma_load_2d(sA + Cfg::A_HALO_BYTES + tap * Cfg::B_BYTES, &pr->tmB, full, tap * Cin + cb * 64,
                                        tc.nt * BLOCK_N);
                        if (++stage == stages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    continue;
                }
                if (row_shared) {
                    // one 18-row box per (dx, channel block) + the weights of its three vertical taps
                    for (int dxi = 0; dxi < 3; ++dxi) {
                        for (int cb = 0; cb < cin_blocks; ++cb) {
                            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                            const uint32_t full = bar_full + 8 * stage;
                            mbar_arrive_expect_tx(full, Cfg::A_RS_BOX_BYTES + 3 * Cfg::B_BYTES);
                            const uint32_t sA = s_tiles + stage * stage_bytes;
                            tma_load_4d(sA, &pr->tmA[0], full, cb * 64, tc.x0 + dxi - 1, tc.y0 - 1, tc.n0);
#pragma unroll
                            for (int dyi = 0; dyi < 3; ++dyi)
                                tma_load_2d(sA + Cfg::A_RS_BYTES + dyi * Cfg::B_BYTES, &pr->tmB, full,
                                            (dyi * 3 + dxi) * Cin + cb * 64, tc.nt * BLOCK_N);
                            if (++stage == stages) {
                                stage = 0;
                                phase ^= 1;
                            }
                        }
                    }
                    continue;
                }
                for (int tap = 0; tap < num_taps; ++tap) {
                    const CUtensorMap* mA = &pr->tmA[p.tap_view[tap]];
                    const int cx = tc.x0 + p.tap_dx[tap], cy = tc.y0 + p.tap_dy[tap];
                    for (int cb = 0; cb < cin_blocks; ++cb) {
                        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                        const uint32_t full = bar_full + 8 * stage;
                        mbar_arrive_expect_tx(full, Cfg::STAGE_BYTES);
                        const uint32_t sA = s_tiles + stage * stage_bytes;
                        tma_load_4d(sA, mA, full, cb * 64, cx, cy, tc.n0);
                        tma_load_2d(sA + Cfg::A_BYTES, &pr->tmB, full, tap * Cin + cb * 64, tc.nt * BLOCK_N);
                        if (++stage == stages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        if (DAFNE_ONE_THREAD(1, lane)) {
            constexpr uint32_t idesc = umma_idesc_f16(128, BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            int g = 0;
            long long bkey = -1;
            uint32_t bphase = 0;
            int a5s = 0;  // mode 5: A ring position
            uint32_t a5ph = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                while (t >= s_begin[g + 1]) ++g;
                const int num_kb =
                    (row_shared >= 2 ? 1 : (row_shared ? 3 : probs[g].p.num_taps)) * probs[g].p.cin_blocks;
                if (row_shared == 3) {
                    const long long key = (static_cast<long long>(reinterpret_cast<uintptr_t>(probs[g].p.w_id)) << 8) |
                                          ((t - s_begin[g]) % probs[g].p.n_tiles);
                    if (key != bkey) {
                        mbar_wait(bar_bfull, bphase);
                        bphase ^= 1;
                        bkey = key;
                    }
                }
                mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + acc * BLOCK_N;
                if (halo5) {
                    const bool gn_in = probs[g].p.in_gn_sums != nullptr;
                    const int cbs = probs[g].p.cin_blocks;
                    for (int cb = 0; cb < cbs; ++cb) {
                        mbar_wait((gn_in ? bar_aready : bar_afull) + 8 * a5s, a5ph);
                        tc_fence_after();
                        const uint64_t a_base = umma_desc_sw128_ex(s_bres + a5s * Cfg::A_HALO_BYTES, 1280, 0);
#pragma unroll 1
                        for (int tap = 0; tap < 9; ++tap) {
                            mbar_wait(bar_full + 8 * stage, phase);
                            tc_fence_after();
                            const uint64_t ad = a_base + (((tap / 3) * 10 + (tap % 3)) * 128 >> 4);
                            const uint64_t bd = umma_desc_sw128(s_tiles + stage * stage_bytes);
#pragma unroll
                            for (int k = 0; k < 4; ++k) umma_f16(d, ad + 2 * k, bd + 2 * k, idesc, (cb | tap | k) != 0);
                            umma_commit(bar_empty + 8 * stage);
                            if (++stage == stages) {
                                stage = 0;
                                phase ^= 1;
                            }
                        }
                        umma_commit(bar_aempty + 8 * a5s);  // the box may be overwritten once its 36 MMAs are done
                        if (++a5s == Cfg::A5_SLOTS) {
                            a5s = 0;
                            a5ph ^= 1;
                        }
                    }
                } else
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(bar_full + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t sA = s_tiles + stage * stage_bytes;
                    if (row_shared >= 2) {
                        // The issuing thread is the bottleneck of these narrow-N convolutions (a 128 x 16 x 16 MMA
                        // is done in a few dozen cycles): both base descriptors are built once per stage and the 36
                        // MMAs differ by compile-time constants only. In the halo modes kb is the channel block.
                        const uint64_t a_base = umma_desc_sw128_ex(sA, 1280, 0);
                        const uint64_t b_base = umma_desc_sw128(row_shared == 3 ? s_bres + kb * 9 * Cfg::B_BYTES
                                                                                : sA + Cfg::A_HALO_BYTES);
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            const uint64_t ad = a_base + (((tap / 3) * 10 + (tap % 3)) * 128 >> 4);
                            const uint64_t bd = b_base + (tap * Cfg::B_BYTES >> 4);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_f16(d, ad + 2 * k, bd + 2 * k, idesc, (tap | k) != 0 ? 1u : (kb != 0));
                        }
                    } else if (row_shared) {
#pragma unroll
                        for (int dyi = 0; dyi < 3; ++dyi) {
                            // rows dyi .. dyi+15 of the 18-row box: the start moves by whole 1024-byte atoms
                            const uint64_t ad = umma_desc_sw128(sA + dyi * 1024);
                            const uint64_t bd = umma_desc_sw128(sA + Cfg::A_RS_BYTES + dyi * Cfg::B_BYTES);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_f16(d, ad + 2 * k, bd + 2 * k, idesc, (kb | dyi | k) != 0);
                        }
                    } else {
                        const uint64_t ad = umma_desc_sw128(sA);
                        const uint64_t bd = umma_desc_sw128(sA + Cfg::A_BYTES);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            // +32 B along K inside the 128 B swizzle row = +2 in the (addr >> 4) field
                            umma_f16(d, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
                        }
                    }
                    umma_commit(bar_empty + 8 * stage);
                    if (++stage == stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(bar_tfull + 8 * acc);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else if ((warp == 2 || warp == 3) && halo5) {
        // ------------------------------------------------------------ GroupNorm + ReLU of the INPUT, on the landed box
        // (mode 5, tower layers 2-4): the previous layer stored its RAW convolution output and its statistics; instead
        // of a separate read-modify-write pass over that tensor (2 x 512 B per location and layer, 1.4 ms per forward at
        // 32 x 1024^2), the two spare warps rewrite each 18 x 10 pixel x 64 channel box in shared memory between the TMA
        // load and the MMAs: y = max(x * a + b, 0) with a = rstd * gamma, b = beta - mean * a (fp32, rounded to fp16 once,
        // like the separate pass). Every element is touched ONCE per tile (the box serves all nine taps) -- 46 KB of shared-memory
        // traffic per channel block next to the 432 KB the nine taps' MMAs read. Pixels outside the image are the
        // convolution's zero padding (TMA zero fill) and stay zero: the padding applies to the NORMALISED map.
        const int tid = threadIdx.x - 64;  // 0..63
        // [2][32] (mean, rstd) of the image's 32 groups (8 channels each), double-buffered over tiles: with a grid
        // stride of 148 and 128 tiles per P3 image nearly every tile of a CTA belongs to another image
        float2* gstat = reinterpret_cast<float2*>(smem + Cfg::A5_RING_BYTES);
        // Thread tid takes sixteen-byte unit qs = tid & 7 of pixels px = (tid >> 3) + 8 k, k < 22.5. With the
        // address-based 128-byte swizzle that unit holds channels 8 * (qs ^ (px & 7)) ... + 7, and px & 7 = (tid >> 3) & 7
        // for every k: the thread always meets the SAME eight channels of a channel block -- exactly one GroupNorm
        // group -- whose (a, b) pairs it keeps in registers for the whole box.
        const int qs = tid & 7, pxb = tid >> 3;
        const int grp_in_cb = qs ^ (pxb & 7);
        int as = 0;
        uint32_t aph = 0;
        int g = 0;
        int buf = 0;
        // raw statistics of this thread's group (tid < 32) for the tile it is about to work on, fetched one tile ahead
        long long pre1 = 0, pre2 = 0;
        auto fetch_stats = [&](int t, int g_from) {
            if (t >= total_tiles || tid >= 32) return;
            int gg = g_from;
            while (t >= s_begin[gg + 1]) ++gg;
            const ConvParams& q = probs[gg].p;
            if (q.in_gn_sums == nullptr || tid >= (q.Cin >> 3)) return;
            const int n0 = tile_coord(q, t - s_begin[gg]).n0;
            const long long* sp = q.in_gn_sums + (static_cast<size_t>(n0) * (q.Cin >> 3) + tid) * 2;
            pre1 = __ldg(sp);
            pre2 = __ldg(sp + 1);
        };
        fetch_stats(blockIdx.x, 0);
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            while (t >= s_begin[g + 1]) ++g;
            const ConvParams& p = probs[g].p;
            const int cbs = p.cin_blocks;
            if (p.in_gn_sums == nullptr) {  // plain input: the MMA warp waits for the TMA itself
                for (int cb = 0; cb < cbs; ++cb)
                    if (++as == Cfg::A5_SLOTS) {
                        as = 0;
                        aph ^= 1;
                    }
                fetch_stats(t + gridDim.x, g);
                continue;
            }
            const TileCoord tc = tile_coord(p, t - s_begin[g]);
            const int Wout = p.Wout, Hout = p.Hout;
            const float* gamma = p.in_gamma;
            const float* beta = p.in_beta;
            if (tid < (p.Cin >> 3)) {
                const float inv_cnt = 1.0f / (static_cast<float>(Hout * Wout) * 8.0f);
                const float s1 = static_cast<float>(static_cast<double>(pre1) * (1.0 / kGnSumScale));
                const float s2 = static_cast<float>(static_cast<double>(pre2) * (1.0 / kGnSqScale));
                const float mean = s1 * inv_cnt;
                const float var = fmaxf(s2 * inv_cnt - mean * mean, 0.f);
                gstat[buf * 32 + tid] = make_float2(mean, rsqrtf(var + 1e-5f));
            }
            fetch_stats(t + gridDim.x, g);
            // which of this thread's 23 pixels lie inside the image (the others are the convolution's zero padding and
            // stay zero): once per tile, the four channel blocks share it
            uint32_t on = 0;
#pragma unroll 1
            for (int k = 0; k < 23; ++k) {
                const int px = pxb + 8 * k;
                const int x = tc.x0 - 1 + px % 10, y = tc.y0 - 1 + px / 10;
                if (px < 180 && x >= 0 && y >= 0 && x < Wout && y < Hout) on |= 1u << k;
            }
            named_bar_sync(3, 64);  // the table of this tile is complete (the other buffer belongs to the previous tile)
            for (int cb = 0; cb < cbs; ++cb) {
                // y = max(x * a + b, 0) with a = rstd * gamma, b = beta - mean * a: one FMA per element
                float2 ab[8];
                {
                    const int grp = cb * 8 + grp_in_cb;
                    const float2 mr = gstat[buf * 32 + grp];
                    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + grp * 8));
                    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + grp * 8) + 1);
                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + grp * 8));
                    const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + grp * 8) + 1);
                    const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                    const float bt[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float a = mr.y * gm[j];
                        ab[j] = make_float2(a, bt[j] - mr.x * a);
                    }
                }
                mbar_wait(bar_afull + 8 * as, aph);
                const uint32_t unit0 = s_bres + as * Cfg::A_HALO_BYTES + pxb * 128 + qs * 16;
                uint32_t m = on;
#pragma unroll 1
                for (int k0 = 0; k0 < 23; k0 += 4, m >>= 4) {  // four units in flight
                    uint32_t w[4][4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (m & (1u << i))
                            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                         : "=r"(w[i][0]), "=r"(w[i][1]), "=r"(w[i][2]), "=r"(w[i][3])
                                         : "r"(unit0 + (k0 + i) * 1024));
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (!(m & (1u << i))) continue;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i][j]));
                            const float a0 = fmaf(f.x, ab[2 * j].x, ab[2 * j].y);
                            const float a1 = fmaf(f.y, ab[2 * j + 1].x, ab[2 * j + 1].y);
                            asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(w[i][j]) : "f"(a1), "f"(a0));
                        }
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(unit0 + (k0 + i) * 1024),
                                     "r"(w[i][0]), "r"(w[i][1]), "r"(w[i][2]), "r"(w[i][3])
                                     : "memory");
                    }
                }
                fence_proxy_async_smem();  // the rewritten box is read by the tensor core (async proxy)
                named_bar_sync(3, 64);
                if (tid == 0) mbar_arrive(bar_aready + 8 * as);
                if (++as == Cfg::A5_SLOTS) {
                    as = 0;
                    aph ^= 1;
                }
            }
            buf ^= 1;
        }
    } else if (warp == 3) {
        // ------------------------------------------------------------ residual producer
        if (BLOCK_N >= 64 && res_tma && DAFNE_ONE_THREAD(4, lane)) {
            constexpr int CHUNKS = BLOCK_N >= 64 ? BLOCK_N / 64 : 1;
            int cnt[2] = {0, 0};  // chunks issued per epilogue warpgroup
            int g = 0, it = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
                while (t >= s_begin[g + 1]) ++g;
                const ConvProblem* pr = probs + g;
                const TileCoord tc = tile_coord(pr->p, t - s_begin[g]);
                const int wg = EPI_WGS == 2 ? (it & 1) : 0;
                for (int j = 0; j < CHUNKS; ++j) {
                    const int c = cnt[wg]++;
                    const int slot = c % ring;
                    const uint32_t round = static_cast<uint32_t>(c / ring);
                    mbar_wait(bar_rempty + 8 * (wg * RS + slot), (round & 1) ^ 1);
                    const uint32_t full = bar_rfull + 8 * (wg * RS + slot);
                    mbar_arrive_expect_tx(full, Cfg::SLOT_BYTES);
                    tma_load_4d(s_epi + (wg * ring + slot) * Cfg::SLOT_BYTES, &pr->tmRes, full,
                                tc.nt * BLOCK_N + j * 64, tc.x0, tc.y0, tc.n0);
                }
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------ epilogue
        if constexpr (EPI_WGS == 2) asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
        const int wg = (warp - 4) >> 2;  // epilogue warpgroup
        const int wi = warp & 3;         // this warp may touch TMEM lanes [32*wi, 32*wi+32)
        const int et = (threadIdx.x - 128) & 127;
        const int row = wi * 32 + lane;
        const uint32_t bar_id = 1 + wg;
        float2* s_tab = s_tab_all + wg * BLOCK_N;  // MODE 2: BLOCK_N floats (shift only) at the same place
        float* s_bias = reinterpret_cast<float*>(s_tab_all) + wg * BLOCK_N;
        const uint32_t s_epi_wg = s_epi + wg * ring * Cfg::SLOT_BYTES;
        uint32_t acc_phase = 0;  // EPI_WGS == 2: this warpgroup always drains accumulator stage `wg`
        int acc = EPI_WGS == 2 ? wg : 0;
        int chunk_cnt = 0;  // chunks this warpgroup has staged so far (ring position)
        int g = 0;
        int table_key = -1;  // (problem, n-tile) the scale/shift table in shared memory belongs to
        int it = 0;          // running tile count of this CTA
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
            if (EPI_WGS == 2 && (it & 1) != wg) continue;
            while (t >= s_begin[g + 1]) ++g;
            const ConvProblem* pr = probs + g;
            // this problem's parameters, in registers for the tile
            struct {
                int N, Hout, Wout, Cout, relu, res_H, res_W, res_shift, out_ld;
                const float *scale, *shift;
                const __half* residual;
                long long* gn_sums;
                float* out_f32;
            } p;
            p.N = pr->p.N;
            p.Hout = pr->p.Hout;
            p.Wout = pr->p.Wout;
            p.Cout = pr->p.Cout;
            p.relu = pr->p.relu;
            p.res_H = pr->p.res_H;
            p.res_W = pr->p.res_W;
            p.res_shift = pr->p.res_shift;
            p.out_ld = pr->p.out_ld;
            p.scale = pr->p.scale;
            p.shift = pr->p.shift;
            p.residual = pr->p.residual;
            p.gn_sums = pr->p.gn_sums;
            p.out_f32 = pr->p.out_f32;
            const int tw = pr->p.tw, th = pr->p.th;
            const int rx = row % tw;
            const int ry = (row / tw) % th;
            const int rn = row / (tw * th);
            const TileCoord tc = tile_coord(pr->p, t - s_begin[g]);
            const int nt = tc.nt, x0 = tc.x0, y0 = tc.y0, n0 = tc.n0;
            const int x = x0 + rx, y = y0 + ry, n = n0 + rn;
            const bool valid = x < p.Wout && y < p.Hout && n < p.N;
            // residual read by the threads themselves: only the FPN top-down add (nearest-upsampled, res_shift == 1)
            const bool res_ldg = MODE == 1 && BLOCK_N >= 64 && p.residual != nullptr && !res_tma;

            const __half* res_row = nullptr;
            if (res_ldg)
                res_row = p.residual +
                          ((static_cast<size_t>(n) * p.res_H + (y >> p.res_shift)) * p.res_W + (x >> p.res_shift)) *
                              p.Cout +
                          nt * BLOCK_N;

            if (table_key != g * 4096 + nt) {
                table_key = g * 4096 + nt;
                named_bar_sync(bar_id, 128);  // everyone is done reading the previous table
                for (int i = et; i < BLOCK_N; i += 128) {
                    const int ch = nt * BLOCK_N + i;
                    float2 v;
                    v.x = (p.scale != nullptr && ch < p.Cout) ? __ldg(p.scale + ch) : 1.0f;
                    v.y = (p.shift != nullptr && ch < p.Cout) ? __ldg(p.shift + ch) : 0.0f;
                    if (MODE == 2)
                        s_bias[i] = v.y;
                    else
                        s_tab[i] = v;
                }
                named_bar_sync(bar_id, 128);
            }

            mbar_wait(bar_tfull + 8 * acc, acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(wi * 32) << 16) + acc * BLOCK_N;

            if constexpr (BLOCK_N >= 64) {
                constexpr int CHUNKS = BLOCK_N / 64;
#pragma unroll 1
                for (int j = 0; j < CHUNKS; ++j, ++chunk_cnt) {
                    const int chbase = j * 64;
                    const int slot = chunk_cnt % ring;
                    const uint32_t buf = s_epi_wg + slot * Cfg::SLOT_BYTES;
                    if (res_tma) {
                        // the residual chunk of this tile, TMA-loaded into the slot the output is staged in
                        mbar_wait(bar_rfull + 8 * (wg * RS + slot), static_cast<uint32_t>(chunk_cnt / ring) & 1);
                    }
                    uint32_t packed[32];
                    if (MODE == 2) {
                        uint32_t v[64];
                        DAFNE_TMEM_LD_X32(taddr + chbase, v);
                        DAFNE_TMEM_LD_X32(taddr + chbase + 32, (v + 32));
                        tmem_ld_wait();
                        if (j == CHUNKS - 1) {
                            // all TMEM reads of this tile are done: hand the accumulator back to the MMA warp
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
                        }
                        epilogue_chunk_bias(v, s_bias + chbase, packed);
                    } else {
                        // two halves of 32 channels, each: accumulators out of TMEM, residual in, math, staged out --
                        // short live ranges (see epilogue_half_math)
#pragma unroll 1
                        for (int h = 0; h < 2; ++h) {
                            uint32_t vh[32];
                            DAFNE_TMEM_LD_X32(taddr + chbase + 32 * h, vh);
                            uint4 res[4];
                            if (res_ldg) {
#pragma unroll
                                for (int q = 0; q < 4; ++q)
                                    res[q] = valid ? __ldg(reinterpret_cast<const uint4*>(res_row + chbase) + h * 4 + q)
                                                   : make_uint4(0, 0, 0, 0);
                            }
                            if (res_tma) {
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    const uint32_t src = buf + row * 128 + (((h * 4 + q) ^ (row & 7)) << 4);
                                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                                 : "=r"(res[q].x), "=r"(res[q].y), "=r"(res[q].z), "=r"(res[q].w)
                                                 : "r"(src)
                                                 : "memory");
                                }
                            }
                            tmem_ld_wait();
                            if (j == CHUNKS - 1 && h == 1) {
                                // all TMEM reads of this tile are done: hand the accumulator back to the MMA warp
                                tc_fence_before();
                                __syncwarp();
                                if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
                            }
                            uint32_t ph[16];
                            const float2* tb = s_tab + chbase + 32 * h;
                            if (MODE == 1) {
                                if (p.relu)
                                    epilogue_half_math<true, true>(vh, tb, res, ph);
                                else
                                    epilogue_half_math<false, true>(vh, tb, res, ph);
                            } else {
                                if (p.relu)
                                    epilogue_half_math<true, false>(vh, tb, res, ph);
                                else
                                    epilogue_half_math<false, false>(vh, tb, res, ph);
                            }
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const uint32_t dst = buf + row * 128 + (((h * 4 + q) ^ (row & 7)) << 4);
                                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(ph[4 * q]),
                                             "r"(ph[4 * q + 1]), "r"(ph[4 * q + 2]), "r"(ph[4 * q + 3])
                                             : "memory");
                            }
                        }
                    }
                    float gs[16];
                    if (MODE == 2) {
                        // statistics of what the next layer will read: the fp16-rounded values
#pragma unroll
                        for (int i = 0; i < 16; ++i) gs[i] = 0.0f;
#pragma unroll
                        for (int c = 0; c < 64; c += 2) {
                            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&packed[c >> 1]));
                            gs[(c >> 3) * 2] += f.x + f.y;
                            gs[(c >> 3) * 2 + 1] += f.x * f.x + f.y * f.y;
                        }
                    }
                    if (MODE == 2 && p.gn_sums != nullptr) {
                        // Per-pixel partials (fixed 8-channel order) become 64-bit fixed point BEFORE any cross-thread
                        // reduction: integer addition is associative, so the statistics are bit-identical from run
                        // to run and independent of tiling / batch composition (no float atomics).
                        const int groups = p.Cout >> 3;
                        const int n_lo = __shfl_sync(0xffffffffu, n, 0);
                        int n_hi = __shfl_sync(0xffffffffu, n, 31);
                        if (n_hi > p.N - 1) n_hi = p.N - 1;
                        long long q[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            q[i] = __float2ll_rn(gs[i] * ((i & 1) ? kGnSqScale : kGnSumScale));
                        for (int img = n_lo; img <= n_hi; ++img) {
                            const bool mine = valid && n == img;
                            long long mv[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i) mv[i] = mine ? q[i] : 0ll;
                            const long long tot = warp_reduce16_scatter(mv, lane);
                            if ((lane & 1) == 0) {
                                const int idx = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 +
                                                ((lane >> 1) & 1);
                                const int gg = (nt * BLOCK_N + chbase) / 8 + (idx >> 1);
                                atomicAdd(reinterpret_cast<unsigned long long*>(p.gn_sums) +
                                              (static_cast<size_t>(img) * groups + gg) * 2 + (idx & 1),
                                          static_cast<unsigned long long>(tot));
                            }
                        }
                    }
                    // Stage the chunk in its ring slot (swizzled like the TMA box), then one TMA store per
                    // 128 px x 64 ch chunk. Every thread writes only the row it (in residual mode) just read, and the
                    // slot is known to be free: in residual mode because its reload was gated on the previous store
                    // (bar_rempty), otherwise because the elected thread waited for that store before the barrier
                    // of the PREVIOUS chunk (wait_group.read ring-2 below).
                    if (MODE == 2) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const uint32_t dst = buf + row * 128 + ((q ^ (row & 7)) << 4);
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(packed[4 * q]),
                                         "r"(packed[4 * q + 1]), "r"(packed[4 * q + 2]), "r"(packed[4 * q + 3])
                                         : "memory");
                        }
                    }
                    fence_proxy_async_smem();
                    if (et == 0 && !res_tma && ring >= 2) {
                        // before anyone may write the NEXT chunk's slot, the store that last read it must be done
                        if (ring == 2)
                            tma_store_wait_read<0>();
                        else if (ring == 3)
                            tma_store_wait_read<1>();
                        else
                            tma_store_wait_read<2>();
                    }
                    named_bar_sync(bar_id, 128);
                    if (et == 0) {
                        tma_store_4d(&pr->tmOut, buf, nt * BLOCK_N + chbase, x0, y0, n0);
                        tma_store_commit();
                        if (res_tma && chunk_cnt > 0) {
                            // the previous chunk's store has read its slot: the residual producer may refill it
                            tma_store_wait_read<1>();
                            mbar_arrive(bar_rempty + 8 * (wg * RS + (chunk_cnt - 1) % ring));
                        }
                    }
                    if (ring == 1 && !res_tma) {
                        // one slot (the tower convolutions trade the second one for a fourth weight stage): the next
                        // chunk goes into THIS slot, so its store must have read it before anyone writes again
                        if (et == 0) tma_store_wait_read<0>();
                        named_bar_sync(bar_id, 128);
                    }
                }
            } else {
                // small-Cout prediction convs: fp32 NHWC rows written straight from registers
                uint32_t v[BLOCK_N];
                if constexpr (BLOCK_N == 16) {
                    DAFNE_TMEM_LD_X16(taddr, v);
                } else {
                    DAFNE_TMEM_LD_X32(taddr, v);
                }
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
                if (valid) {
                    float* op = p.out_f32 + ((static_cast<size_t>(n) * p.Hout + y) * p.Wout + x) * p.out_ld;
#pragma unroll
                    for (int c = 0; c < BLOCK_N; c += 4) {
                        if (c < p.out_ld) {
                            const float4 t0 = *reinterpret_cast<const float4*>(&s_tab[c]);
                            const float4 t1 = *reinterpret_cast<const float4*>(&s_tab[c + 2]);
                            float4 o;
                            o.x = fmaf(__uint_as_float(v[c]), t0.x, t0.y);
                            o.y = fmaf(__uint_as_float(v[c + 1]), t0.z, t0.w);
                            o.z = fmaf(__uint_as_float(v[c + 2]), t1.x, t1.y);
                            o.w = fmaf(__uint_as_float(v[c + 3]), t1.z, t1.w);
                            if (p.relu) {
                                o.x = fmaxf(o.x, 0.f);
                                o.y = fmaxf(o.y, 0.f);
                                o.z = fmaxf(o.z, 0.f);
                                o.w = fmaxf(o.w, 0.f);
                            }
                            *reinterpret_cast<float4*>(op + c) = o;
                        }
                    }
                }
            }
            if (EPI_WGS == 2) {
                acc_phase ^= 1;
            } else {
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
        if (BLOCK_N >= 64 && et == 0) tma_store_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------ host side

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (fn) return fn;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || ptr == nullptr) {
        set_error("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %s", cudaGetErrorString(e));
        return nullptr;
    }
    fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    return fn;
}

int encode_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
               const uint32_t* box, CUtensorMapL2promotion promo, const char* what) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return -1;
    cuuint64_t gdim[5];
    cuuint64_t gstr[4];
    cuuint32_t bdim[5];
    cuuint32_t estr[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estr[i] = 1;
    }
    for (int i = 0; i < rank - 1; ++i) gstr[i] = strides_bytes[i];
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(%s) failed with CUresult %d (dims %llu %llu %llu %llu box %u %u %u %u)", what,
                  (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1],
                  (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
                  box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
        return -1;
    }
    return 0;
}

static int pow2ceil(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

int conv_plan_build(const ConvDesc& d, ConvPlan* plan, int num_sms) {
    memset(plan, 0, sizeof(*plan));
    if (d.Cin % 64 != 0) {
        set_error("conv_tc: Cin=%d must be a multiple of 64", d.Cin);
        return -1;
    }
    if (!(d.ksize == 1 || d.ksize == 3) || !(d.stride == 1 || d.stride == 2)) {
        set_error("conv_tc: unsupported ksize=%d stride=%d", d.ksize, d.stride);
        return -1;
    }
    const bool small = d.out_f32 != nullptr;
    int bn;
    if (small) {
        if (d.Cout > 32 || d.out_ld % 4 != 0 || d.out_ld < d.Cout) {
            set_error("conv_tc: fp32-output path needs Cout<=32 and out_ld%%4==0 (Cout=%d ld=%d)", d.Cout, d.out_ld);
            return -1;
        }
        bn = d.Cout <= 16 ? 16 : 32;
    } else {
        if (d.Cout % 64 != 0) {
            set_error("conv_tc: Cout=%d must be a multiple of 64", d.Cout);
            return -1;
        }
        bn = d.Cout % 256 == 0 ? 256 : (d.Cout % 128 == 0 ? 128 : 64);
        // The bottleneck-closing 1x1 convolutions of res2-res4 (K <= 256, residual at the output's resolution) are
        // HBM-bound and their critical path is the residual stream through the epilogue ring: 128-wide tiles leave
        // shared memory for a deeper ring and more operand stages (r2n A/B on one B200, per forward: res2 conv3
        // 382 -> 302 us and res3 303 -> 239 us with ring 4 / 3 stages; res4 282 -> 237 us with ring 3 / 4 stages;
        // ring 5 / 2 stages and ring 2 are both worse, the latter by 50 %).
        const bool res_1x1 = d.residual != nullptr && d.res_shift == 0 && d.ksize == 1 && d.Cin <= 256;
        if (res_1x1 && d.Cout % 128 == 0) bn = 128;
        if (const char* ov = getenv("DAFNE_CONV_RES_BN")) {  // tuning aid: tile width of those convolutions
            const int v = atoi(ov);
            if (res_1x1 && (v == 256 || v == 128 || v == 64) && d.Cout % v == 0) bn = v;
        }
    }
    ConvParams& p = plan->prob.p;
    p.N = d.N;
    p.Hout = d.Hout;
    p.Wout = d.Wout;
    p.Cin = d.Cin;
    p.Cout = d.Cout;
    p.num_taps = d.ksize * d.ksize;
    p.cin_blocks = d.Cin / 64;
    // Row-shared taps: worth it where the 9 A loads per channel block are the bottleneck (narrow N tiles); the
    // three weight tiles of a stage must leave room for a four-stage pipeline (bn <= 64).
    const bool row_shared = d.ksize == 3 && d.stride == 1 && bn <= 64;
    plan->row_shared = row_shared ? 1 : 0;
    // 2 = halo box (one A load per channel block serves all nine taps); DAFNE_CONV_TAPS=rows selects the older
    // row-shared form (three loads, atom-aligned descriptor starts only) for A/B measurements
    bool halo = row_shared;
    if (const char* hv = getenv("DAFNE_CONV_TAPS"))
        if (strcmp(hv, "rows") == 0) halo = false;
    if (halo) plan->row_shared = 2;
    plan->breg_bytes = 0;
    if (halo && 9 * (d.Cin / 64) * bn * 128 <= kMaxResidentB) {
        // weights resident in shared memory: measured equal to streaming them per stage (both end up bound by the
        // tensor core's ~76 cycles per 128 x 16 x 16 MMA), so it stays an A/B option
        if (const char* hv = getenv("DAFNE_CONV_TAPS"))
            if (strcmp(hv, "resident") == 0) {
                plan->row_shared = 3;
                plan->breg_bytes = 9 * (d.Cin / 64) * bn * 128;
            }
    }
    // mode 5: the 256-wide tower convolutions (GroupNorm statistics in the epilogue) take their input as halo boxes too
    // -- one A load per channel block instead of nine, 28 % less L2 -> shared-memory traffic -- with the weights of one
    // tap per stage, and can normalise the input while it is loaded (in_gn_sums). DAFNE_CONV_HALO256=0: A/B switch.
    // (Measured and not adopted, r3zg: the same operand path for the plain 256-wide 3x3 convolutions -- conv2 of res4 /
    // res5, the FPN output convolutions: 2.53 / 2.92 / 2.86 ms against 2.53 / 2.96 / 2.76 ms for res4's conv2 in three
    // interleaved rounds, no difference; they already run four (A + B) stages.)
    bool mode5 = !small && d.ksize == 3 && d.stride == 1 && bn == 256 && d.gn_sums != nullptr && d.Cin <= 256;
    if (const char* ev = getenv("DAFNE_CONV_HALO256"))
        if (atoi(ev) == 0) mode5 = false;
    if (d.in_gn_sums != nullptr && !mode5) {
        set_error("conv_tc: GroupNorm on load needs the 256-wide halo mode (3x3, stride 1, Cout %% 256 == 0, Cin <= 256)");
        return -1;
    }
    if (mode5) {
        plan->row_shared = 5;
        plan->breg_bytes = ConvCfg<256, 1>::A5_REGION_BYTES;
        halo = true;
    }
    if (row_shared || mode5) {
        p.tw = 8;
        p.th = 16;
        p.nb = 1;
    } else {
        p.tw = pow2ceil(d.Wout) < 16 ? pow2ceil(d.Wout) : 16;
        p.th = pow2ceil(d.Hout) < 128 / p.tw ? pow2ceil(d.Hout) : 128 / p.tw;
        p.nb = 128 / (p.tw * p.th);
    }
    p.tiles_x = (d.Wout + p.tw - 1) / p.tw;
    p.tiles_y = (d.Hout + p.th - 1) / p.th;
    p.tiles_n = (d.N + p.nb - 1) / p.nb;
    p.m_tiles = p.tiles_x * p.tiles_y * p.tiles_n;
    p.n_tiles = (d.Cout + bn - 1) / bn;
    p.total_tiles = p.m_tiles * p.n_tiles;
    p.tile_begin = 0;
    p.reverse_m = d.reverse_m;
    p.scale = d.scale;
    p.shift = d.shift;
    p.relu = d.relu;
    p.residual = d.residual;
    p.res_H = d.res_H;
    p.res_W = d.res_W;
    p.res_shift = d.res_shift;
    p.gn_sums = d.gn_sums;
    p.in_gn_sums = d.in_gn_sums;
    p.in_gamma = d.in_gamma;
    p.in_beta = d.in_beta;
    p.out_f32 = d.out_f32;
    p.out_ld = d.out_ld;
    p.w_id = d.w;
    plan->block_n = bn;
    plan->grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
    plan->flops = 2.0 * d.N * d.Hout * d.Wout * (double)d.Cout * p.num_taps * d.Cin;

    const uint64_t C = d.Cin, W = d.Win, H = d.Hin;
    const uint32_t boxA[4] = {64u, (uint32_t)p.tw, (uint32_t)p.th, (uint32_t)p.nb};
    const uint32_t boxA_rs[4] = {64u, halo ? 10u : 8u, 18u, 1u};  // the tile plus one row above and below
    bool view_empty[4] = {false, false, false, false};
    if (d.stride == 1) {
        const uint64_t dims[4] = {C, W, H, (uint64_t)d.N};
        const uint64_t str[3] = {C * 2, W * C * 2, H * W * C * 2};
        if (encode_map(&plan->prob.tmA[0], d.in, 4, dims, str, (row_shared || mode5) ? boxA_rs : boxA,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "A"))
            return -1;
        for (int v = 1; v < 4; ++v) plan->prob.tmA[v] = plan->prob.tmA[0];
        for (int t = 0; t < p.num_taps; ++t) {
            p.tap_view[t] = 0;
            p.tap_dy[t] = d.ksize == 3 ? t / 3 - 1 : 0;
            p.tap_dx[t] = d.ksize == 3 ? t % 3 - 1 : 0;
        }
    } else {
        // stride 2: four parity views (rows py, py+2, ...; cols px, px+2, ...) of the same tensor; every tap of the
        // 3x3 (or the single 1x1 tap) becomes a unit-stride box in one of them.
        for (int py = 0; py < 2; ++py)
            for (int px = 0; px < 2; ++px) {
                const int v = py * 2 + px;
                uint64_t vw = (W - px + 1) / 2, vh = (H - py + 1) / 2;
                if ((int64_t)W - px <= 0) vw = 0;
                if ((int64_t)H - py <= 0) vh = 0;
                if (vw == 0 || vh == 0) {
                    view_empty[v] = true;
                    vw = vw ? vw : 1;
                    vh = vh ? vh : 1;
                }
                const uint64_t dims[4] = {C, vw, vh, (uint64_t)d.N};
                const uint64_t str[3] = {2 * C * 2, 2 * W * C * 2, H * W * C * 2};
                const __half* base = d.in + ((size_t)py * W + px) * C;
                if (view_empty[v]) base = d.in;
                if (encode_map(&plan->prob.tmA[v], base, 4, dims, str, boxA, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "A-s2"))
                    return -1;
            }
        for (int t = 0; t < p.num_taps; ++t) {
            const int ky = d.ksize == 3 ? t / 3 : 1, kx = d.ksize == 3 ? t % 3 : 1;
            const int py = ky == 1 ? 0 : 1, px = kx == 1 ? 0 : 1;
            p.tap_view[t] = py * 2 + px;
            p.tap_dy[t] = ky == 0 ? -1 : 0;
            p.tap_dx[t] = kx == 0 ? -1 : 0;
            if (view_empty[p.tap_view[t]]) p.tap_dx[t] = 1 << 20;  // whole box out of bounds -> zeros
        }
    }
    {
        const uint64_t K = (uint64_t)p.num_taps * d.Cin;
        const uint64_t dims[2] = {K, (uint64_t)d.Cout};
        const uint64_t str[1] = {K * 2};
        const uint32_t box[2] = {64u, (uint32_t)bn};
        if (encode_map(&plan->prob.tmB, d.w, 2, dims, str, box, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "B")) return -1;
    }
    if (!small) {
        const uint64_t Co = d.Cout, Wo = d.Wout, Ho = d.Hout;
        const uint64_t dims[4] = {Co, Wo, Ho, (uint64_t)d.N};
        const uint64_t str[3] = {Co * 2, Wo * Co * 2, Ho * Wo * Co * 2};
        if (encode_map(&plan->prob.tmOut, d.out, 4, dims, str, boxA, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "Out")) return -1;
    } else {
        plan->prob.tmOut = plan->prob.tmB;
    }
    plan->prob.tmRes = plan->prob.tmOut;
    plan->res_tma = 0;
    plan->mode = d.residual != nullptr ? 1 : (d.gn_sums != nullptr ? 2 : 0);
    if (d.gn_sums != nullptr && (d.residual != nullptr || d.scale != nullptr || d.relu)) {
        set_error("conv_tc: GroupNorm statistics go with a bias-only convolution (no scale, residual or ReLU)");
        return -1;
    }
    if (!small && d.residual != nullptr && d.res_shift == 0) {
        if (d.res_H != d.Hout || d.res_W != d.Wout) {
            set_error("conv_tc: residual %dx%d does not match the output %dx%d", d.res_H, d.res_W, d.Hout, d.Wout);
            return -1;
        }
        const uint64_t Co = d.Cout, Wo = d.Wout, Ho = d.Hout;
        const uint64_t dims[4] = {Co, Wo, Ho, (uint64_t)d.N};
        const uint64_t str[3] = {Co * 2, Wo * Co * 2, Ho * Wo * Co * 2};
        if (encode_map(&plan->prob.tmRes, d.residual, 4, dims, str, boxA, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "Res"))
            return -1;
        // res_tma carries the ring depth asked for (>= 2): the residual prefetch distance, paid for in operand stages
        plan->res_tma = 4;
        if (d.ksize == 1 && bn == 128 && d.Cin > 128) plan->res_tma = 3;  // K = 256: four operand stages matter more
        if (const char* ov = getenv("DAFNE_CONV_RING")) {  // tuning aid: 2 .. MAX_RING
            const int v = atoi(ov);
            if (v >= 2 && v <= 6) plan->res_tma = v;
        }
    }
    // K <= 256 (every 1x1 of res2-res4, the 3x3 of res2): HBM-bound, the epilogue is the critical path
    plan->epi_wgs = (!small && p.num_taps * d.Cin <= 256) ? 2 : 1;
    if (const char* ov = getenv("DAFNE_CONV_WGS")) {  // tuning aid (scripts/prof_conv.py): force 1 | 2
        const int v = atoi(ov);
        if (!small && (v == 1 || v == 2)) plan->epi_wgs = v;
    }
    return 0;
}

// Operand stages / epilogue ring slots per warpgroup for a launch: as deep as 227 KB allows. With a TMA residual the
// ring is the prefetch depth of the residual stream, so it gets four slots at the price of operand stages.
template <int BN, int WGS>
static void conv_smem_config(int res_tma, int row_shared, int breg, int mode, int* stages, int* ring) {
    using Cfg = ConvCfg<BN, WGS>;
    int r = BN < 64 ? 2 : (res_tma ? (res_tma >= 2 && res_tma <= Cfg::MAX_RING ? res_tma : 4) : (WGS == 2 ? 3 : 2));
    if (row_shared == 5) {
        // tower convolutions: the epilogue has a whole tile's 18 432 tensor cycles to drain 4 chunks; a one-slot output
        // ring buys a fourth weight stage. DAFNE_CONV_RING5: tuning override.
        r = 1;
        if (const char* ev = getenv("DAFNE_CONV_RING5")) r = atoi(ev) >= 1 && atoi(ev) <= 4 ? atoi(ev) : r;
    }
    int st = Cfg::MAX_STAGES;
    while (st > 2 && Cfg::smem_bytes(st, r, row_shared, breg, mode) > kMaxSmem) --st;
    while (r > 2 && Cfg::smem_bytes(st, r, row_shared, breg, mode) > kMaxSmem) --r;
    *stages = st;
    *ring = r;
}

template <int BN, int WGS, int MODE>
static int launch_bn(const ConvProblem* dev_probs, int nprob, int total_tiles, int grid, int res_tma, int row_shared,
                     int breg, cudaStream_t stream) {
    using Cfg = ConvCfg<BN, WGS>;
    static DeviceOnce configured;
    int dev;
    if (!configured.get(&dev)) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN, WGS, MODE>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(conv_tc_kernel<%d,%d,%d>, smem=%d): %s", BN, WGS, MODE, kMaxSmem,
                      cudaGetErrorString(e));
            return -1;
        }
        configured.set(dev, 1);
    }
    int stages, ring;
    conv_smem_config<BN, WGS>(res_tma, row_shared, breg, MODE, &stages, &ring);
    const int smem = Cfg::smem_bytes(stages, ring, row_shared, breg, MODE);
    if (smem > kMaxSmem) {
        set_error("conv_tc_kernel<%d,%d>: %d stages + %d ring slots need %d bytes of shared memory", BN, WGS, stages,
                  ring, smem);
        return -1;
    }
    conv_tc_kernel<BN, WGS, MODE>
        <<<grid, Cfg::THREADS, smem, stream>>>(dev_probs, nprob, total_tiles, stages, ring, res_tma, row_shared, breg);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("conv_tc_kernel<%d,%d,%d> launch: %s", BN, WGS, MODE, cudaGetErrorString(e));
        return -1;
    }
    return 0;
}

template <int BN, int WGS>
static int launch_mode(const ConvProblem* dev_probs, int nprob, int total_tiles, int grid, int mode, int res_tma,
                       int row_shared, int breg, cudaStream_t stream) {
    if (mode == 1) return launch_bn<BN, WGS, 1>(dev_probs, nprob, total_tiles, grid, res_tma, row_shared, breg, stream);
    if (mode == 0) return launch_bn<BN, WGS, 0>(dev_probs, nprob, total_tiles, grid, 0, row_shared, breg, stream);
    set_error("conv_tc: epilogue mode %d is not built for tile width %d / %d epilogue warpgroups", mode, BN, WGS);
    return -1;
}

int conv_group_launch(const ConvProblem* dev_probs, int nprob, int total_tiles, int block_n, int epi_wgs, int mode,
                      int res_tma, int row_shared, int breg, int num_sms, cudaStream_t stream) {
    if (total_tiles == 0) return 0;
    if (nprob < 1 || nprob > kMaxConvProblems) {
        set_error("conv_tc: %d problems in one launch (1..%d supported)", nprob, kMaxConvProblems);
        return -1;
    }
    const int grid = total_tiles < num_sms ? total_tiles : num_sms;
    const int key = block_n * 10 + epi_wgs;
    if (mode == 2) {
        // GroupNorm statistics: only the 256-wide tower convolutions produce them
        if (key == 2561) return launch_bn<256, 1, 2>(dev_probs, nprob, total_tiles, grid, 0, row_shared, breg, stream);
        if (key == 2562) return launch_bn<256, 2, 2>(dev_probs, nprob, total_tiles, grid, 0, 0, 0, stream);
        set_error("conv_tc: GroupNorm statistics need Cout %% 256 == 0 (tile width %d)", block_n);
        return -1;
    }
    switch (key) {
        case 161: return launch_bn<16, 1, 0>(dev_probs, nprob, total_tiles, grid, 0, row_shared, breg, stream);
        case 321: return launch_bn<32, 1, 0>(dev_probs, nprob, total_tiles, grid, 0, row_shared, breg, stream);
        case 641: return launch_mode<64, 1>(dev_probs, nprob, total_tiles, grid, mode, res_tma, row_shared, breg, stream);
        case 642: return launch_mode<64, 2>(dev_probs, nprob, total_tiles, grid, mode, res_tma, row_shared, breg, stream);
        case 1281: return launch_mode<128, 1>(dev_probs, nprob, total_tiles, grid, mode, res_tma, row_shared, 0, stream);
        case 1282: return launch_mode<128, 2>(dev_probs, nprob, total_tiles, grid, mode, res_tma, row_shared, 0, stream);
        case 2561: return launch_mode<256, 1>(dev_probs, nprob, total_tiles, grid, mode, res_tma, 0, 0, stream);
        case 2562: return launch_mode<256, 2>(dev_probs, nprob, total_tiles, grid, mode, res_tma, 0, 0, stream);
    }
    set_error("conv_tc: unsupported tile width %d with %d epilogue warpgroups", block_n, epi_wgs);
    return -1;
}

}  // namespace dafne
