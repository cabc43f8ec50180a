// Greedy polygon NMS on device, lazily evaluated: only the IoUs the sequential sweep can consult are computed.
//
// Semantics (dafne/modeling/nms/nms.py:37-92 -> external poly_gpu_nms): boxes arrive sorted by descending score; box j
// is kept iff no KEPT earlier box i has IoU(i, j) > thr. The result only depends on IoU(i, j) for kept i, so instead of
// the reference's full n x n bitmask (n^2/2 polygon clips, ~10 MB mask, host sweep) the boxes are processed in panels of
// 512:
//   nms_diag_kernel   panel x panel upper triangle for the rows/columns still alive, then (last block done) the
//                     sequential sweep inside the panel -> the panel's kept rows
//   nms_bcast_kernel  kept rows of the panel x every later box still alive; sets `removed` bits (one wave of resident
//                     CTAs pulling (column chunk, row chunk, image) items from a counter). Run TWICE per panel: pass 0
//                     evaluates only the pairs whose centres are close (the likely suppressors), pass 1 the remaining
//                     pairs of the columns that are STILL alive -- a suppressed box needs one hit, and once its near
//                     twin has been found none of its other pairs is clipped. Which pairs go first changes the work,
//                     never a decision: a column dies iff SOME kept row has IoU > thr with it.
// Work drops from n^2/2 pairs to about (kept x alive) pairs. Inside both kernels a pair first goes through
// pair_inter_is_zero() (polyiou.cuh: proves inter == 0 for separated boxes without running the clip); the pairs that
// need the full fp32 clip are compacted into a shared-memory queue and clipped term by term (clip_queue below): a
// queued pair is 16 signed triangle overlaps (edge triangle i of P x edge triangle j of Q, polyiou.cpp:91-103), of which
// only the ones not provably zero are evaluated, one per lane, and added in the reference's order.
// No decision differs from evaluating iou_poly_f32(i, j) > thr for every consulted pair.
#include <stdio.h>
#include <stdlib.h>

#include "conv_tc.cuh"  // set_error
#include "polyiou.cuh"
#include "postprocess.cuh"

namespace dafne {

#ifndef DAFNE_NMS_PANEL
#define DAFNE_NMS_PANEL 512
#endif
constexpr int kPanel = DAFNE_NMS_PANEL;  // <= 512: the sweep keeps the panel's hit words in 48 KB of static shared memory
constexpr int kPanelWords = kPanel / 64;  // 8
constexpr int kDiagBlocks = kPanelWords * (kPanelWords + 1) / 2;  // 36 (rb <= cb)
// Work items are deliberately small (16 x 64 pairs): the clips a CTA ends up with vary by an order of magnitude with
// the local box density, and the hardware's dynamic CTA dispatch balances many small CTAs far better than few big
// ones (32 x 128 tiles left the SMs idle half of the time behind the densest tiles).
constexpr int kRowChunk = 16;  // kept rows per bcast CTA
constexpr int kColChunk = 64;  // columns per bcast CTA
constexpr int kBcastThreads = 256;
constexpr int kBcastImages = 256;  // images one broadcast launch can index (larger batches: one launch per group)
#ifndef DAFNE_DIAG_SPLIT
#define DAFNE_DIAG_SPLIT 4
#endif
// a 64 x 64 diagonal-panel block is worked on by 4 CTAs of 16 rows each (measured r2k, diag kernels per step:
// split 4 = 402 us, 8 = 416 us, 16 = 480 us)
constexpr int kDiagSplit = DAFNE_DIAG_SPLIT;

typedef unsigned long long u64;
// work counters of the last run_nms (diagnostics, bench.py): pairs the sweep consulted and pairs that needed the clip,
// for the diagonal panels and the broadcast separately (slots 2 and 5 are reserved)
constexpr int kNmsStats = 8;
enum { kStDiagPairs = 0, kStDiagQueued = 1, kStBcastPairs = 3, kStBcastQueued = 4 };

static size_t a256n(size_t v) { return (v + 255) / 256 * 256; }

struct NmsLayout {
    size_t o_aux, o_removed, o_diag, o_pk, o_ctr, o_stats, o_work, total;
    int nblk;
};
static NmsLayout nms_layout(int N, int max_sel) {
    NmsLayout y;
    y.nblk = (max_sel + 63) / 64;
    size_t o = 0;
    y.o_aux = o;
    o = a256n(o + static_cast<size_t>(N) * max_sel * sizeof(NmsAux));
    y.o_removed = o;
    o = a256n(o + static_cast<size_t>(N) * y.nblk * 8);
    y.o_pk = o;
    o = a256n(o + static_cast<size_t>(N) * kPanelWords * 8);
    y.o_ctr = o;
    o = a256n(o + static_cast<size_t>(N) * 4);
    y.o_stats = o;
    o = a256n(o + kNmsStats * 8);
    y.o_work = o;  // one work-item counter per (panel, image group) of the broadcast
    o = a256n(o + static_cast<size_t>((max_sel + kPanel - 1) / kPanel) * 2 * ((N + kBcastImages - 1) / kBcastImages) * 4);
    y.o_diag = o;
    o = a256n(o + static_cast<size_t>(N) * kPanel * kPanelWords * 8);
    y.total = o;
    return y;
}
size_t nms_scratch_bytes(int N, int max_sel) { return nms_layout(N, max_sel < 1 ? 1 : max_sel).total; }

// ------------------------------------------------------------------------------------------------ per-box scalars
__global__ void nms_aux_kernel(const float* __restrict__ boxes, const int* __restrict__ counts, int max_sel,
                               NmsAux* __restrict__ aux) {
    const int n = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= counts[n]) return;
    const size_t r = static_cast<size_t>(n) * max_sel + i;
    float b[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) b[k] = boxes[r * 8 + k];
    aux[r] = nms_aux_of(b);
}

// Boxes are staged in shared memory already (re)oriented like iou_poly_f32 does (polyiou.cpp:95-96), once per box.
__device__ __forceinline__ void stage_oriented(const float* __restrict__ src, float* dst) {
    float b[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) b[k] = src[k];
    P2 p[6];
    load_oriented(b, p);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        dst[2 * k] = p[k].x;
        dst[2 * k + 1] = p[k].y;
    }
}

// Appends `want` lanes' entries to a shared-memory queue with one atomic per warp. All 32 lanes must call it.
__device__ __forceinline__ void queue_push(unsigned short* queue, int* qn, bool want, unsigned short entry,
                                           unsigned lane) {
    const unsigned m = __ballot_sync(0xffffffffu, want);
    if (m == 0) return;
    const int leader = __ffs(m) - 1;
    int base = 0;
    if (static_cast<int>(lane) == leader) base = atomicAdd(qn, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (want) queue[base + __popc(m & ((1u << lane) - 1u))] = entry;
}

// Phase 2 of both kernels: the queued pairs are clipped TERM by term. A pair is 16 signed triangle overlaps (edge
// triangle i of P x edge triangle j of Q, polyiou.cpp:91-103); on the pairs that reach this point 8 of the 16 are zero
// on average and polyiou.cuh::term_is_zero proves it for them without clipping. Per batch of kTermBatch pairs:
//   A  one thread per (pair, i): the four terms' zero tests; the live terms go to a shared-memory term queue
//   B  one thread per queued term: tri_overlap -> val[pair][term]       (every lane of every warp holds a live term)
//   C  one thread per pair: the 16 values added in the reference's order (skipped ones are +0: x + 0 == x), the
//      algorithm's own a1 / a2, IoU > thr
// (r1e tried a term queue with the first shortcut alone -- 14 of 16 terms live -- and lost 8 %; with half of the
// terms gone the balance flips.)
constexpr int kTermBatch = 128;
struct TermSmem {
    float val[kTermBatch][17];  // 17: thread `pair` walks its row, rows 17 words apart are conflict-free
    unsigned short tq[kTermBatch * 16];
    int tqn;
};

template <bool BCAST, int THREADS>
__device__ __forceinline__ void clip_queue(const unsigned short* queue, int qn, const float (*rbox)[8],
                                           const float (*cbox)[8], const NmsAux* raux, const NmsAux* caux,
                                           const float (*rt)[4], const float (*ct)[4], TermSmem& ts, float2* poly,
                                           float thr, u64* bits, unsigned int* newdead) {
    const int t = threadIdx.x;
    const unsigned lane = t & 31;
    constexpr int kRowShift = BCAST ? 7 : 6;
    constexpr int kColMask = BCAST ? 127 : 63;
    for (int b0 = 0; b0 < qn; b0 += kTermBatch) {
        const int nb = min(kTermBatch, qn - b0);
        for (int i = t; i < nb * 17; i += THREADS) (&ts.val[0][0])[i] = 0.f;
        if (t == 0) ts.tqn = 0;
        __syncthreads();
        // A: zero tests. Uniform trip count: every lane of a warp reaches queue_push.
        for (int k0 = 0; k0 < nb * 4; k0 += THREADS) {
            const int k = k0 + t;
            bool valid = k < nb * 4;
            const int pl = valid ? (k >> 2) : 0, i = k & 3;
            const int e = queue[b0 + pl];
            const int r = e >> kRowShift, j = e & kColMask;
            // a column some kept row already hit needs no further clip (any hit is enough); atomic read of the 0 -> 1
            // flag so that concurrent warps are race-free by construction
            if (BCAST && valid && atomicOr(&newdead[j], 0u)) valid = false;
            const TermPairCtx ctx = term_pair_ctx(raux[r], caux[j]);
            P2 a, b;
            a.x = rbox[r][2 * i];
            a.y = rbox[r][2 * i + 1];
            b.x = rbox[r][2 * ((i + 1) & 3)];
            b.y = rbox[r][2 * ((i + 1) & 3) + 1];
            const float ta = rt[r][i], tb = rt[r][(i + 1) & 3];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                P2 c, d;
                c.x = cbox[j][2 * jj];
                c.y = cbox[j][2 * jj + 1];
                d.x = cbox[j][2 * ((jj + 1) & 3)];
                d.y = cbox[j][2 * ((jj + 1) & 3) + 1];
                const bool live = valid && !term_is_zero(ctx, a, b, c, d, ta, tb, ct[j][jj], ct[j][(jj + 1) & 3]);
                queue_push(ts.tq, &ts.tqn, live, static_cast<unsigned short>((pl << 4) | (i * 4 + jj)), lane);
            }
        }
        __syncthreads();
        // B: one live term per thread
        const int tn = ts.tqn;
        for (int e0 = 0; e0 < tn; e0 += THREADS) {
            const int en = e0 + t;
            if (en < tn) {
                const int te = ts.tq[en];
                const int pl = te >> 4, term = te & 15, i = term >> 2, jj = term & 3;
                const int e = queue[b0 + pl];
                const int r = e >> kRowShift, j = e & kColMask;
                P2 a, b, c, d;
                a.x = rbox[r][2 * i];
                a.y = rbox[r][2 * i + 1];
                b.x = rbox[r][2 * ((i + 1) & 3)];
                b.y = rbox[r][2 * ((i + 1) & 3) + 1];
                c.x = cbox[j][2 * jj];
                c.y = cbox[j][2 * jj + 1];
                d.x = cbox[j][2 * ((jj + 1) & 3)];
                d.y = cbox[j][2 * ((jj + 1) & 3) + 1];
                ts.val[pl][term] = tri_overlap(a, b, c, d, poly, THREADS);
            }
        }
        __syncthreads();
        // C: the pair's IoU from its 16 terms, in the reference's order (i outer, j inner)
        for (int pl = t; pl < nb; pl += THREADS) {
            float inter = 0.f;
#pragma unroll
            for (int k = 0; k < 16; ++k) inter += ts.val[pl][k];
            const int e = queue[b0 + pl];
            const int r = e >> kRowShift, j = e & kColMask;
            const float uni = raux[r].area + caux[j].area - inter;
            const float iou = (uni == 0.f) ? (inter + 1.f) / (uni + 1.f) : inter / uni;
            if (iou > thr) {
                if (BCAST)
                    atomicExch(&newdead[j], 1u);
                else
                    atomicOr(&bits[r], 1ull << j);
            }
        }
        __syncthreads();  // val / tq are rewritten by the next batch
    }
}

// ------------------------------------------------------------------------------------------------ diagonal panel
constexpr int kDiagThreads = 256;
struct DiagSmem {
    float rbox[64][8];
    float cbox[64][8];
    NmsAux raux[64];
    NmsAux caux[64];
    u64 bits[64];
    unsigned short queue[(64 / kDiagSplit) * 64];
    int qn;
    int mqn;  // pairs the cheap exits of the pre-filter left undecided (listed in terms.tq, which phase 2 reuses)
    int last;
    unsigned stat_pairs;
    float rt[64][4], ct[64][4];     // t(v) of the staged boxes' vertices (term_is_zero)
    float2 poly[9 * kDiagThreads];  // tri_overlap's per-thread polygon columns
    TermSmem terms;
};

// grid (36 * kDiagSplit, N), 256 threads. Block b / kDiagSplit -> (rb, cb), rb <= cb, both 64-box blocks of panel `panel`.
__global__ void __launch_bounds__(kDiagThreads) nms_diag_kernel(const float* __restrict__ boxes, const NmsAux* __restrict__ aux,
                                                      const int* __restrict__ counts, int max_sel, int nblk,
                                                      int panel, float thr, u64* __restrict__ removed,
                                                      u64* __restrict__ diag, u64* __restrict__ pk,
                                                      int* __restrict__ ctr, int* __restrict__ keep,
                                                      int* __restrict__ nkeep, u64* __restrict__ stats) {
    const int n = blockIdx.y;
    const int m = counts[n];
    const int base = panel * kPanel;
    if (base >= m) return;  // uniform for the whole image: nobody counts, nothing to resolve
    // the pair phase's tile and the sweep's copy of the panel's hit words (32 KB) share the same shared memory
    constexpr size_t kSweepBytes = static_cast<size_t>(kPanel) * kPanelWords * sizeof(u64);
    __shared__ __align__(16) unsigned char smem_raw[sizeof(DiagSmem) > kSweepBytes ? sizeof(DiagSmem) : kSweepBytes];
    DiagSmem& sm = *reinterpret_cast<DiagSmem*>(smem_raw);
    // decode (rb, cb) from the linear upper-triangle index; `sub` = which 16 rows of the 64-row block
    const int sub = blockIdx.x % kDiagSplit;
    int rb = 0, rem = blockIdx.x / kDiagSplit;
    while (rem >= kPanelWords - rb) {
        rem -= kPanelWords - rb;
        ++rb;
    }
    const int cb = rb + rem;
    const int t = threadIdx.x;
    const size_t ibase = static_cast<size_t>(n) * max_sel;
    u64* rmv = removed + static_cast<size_t>(n) * nblk;
    const int r0 = base + rb * 64, c0 = base + cb * 64;
    u64 bits_out = 0;
    if (r0 < m && c0 < m) {
        const u64 rdead = rmv[r0 >> 6], cdead = rmv[c0 >> 6];
        const int row = r0 + t, col = c0 + t;
        if (t < 64) {
            if (row < m) {
                stage_oriented(boxes + (ibase + row) * 8, sm.rbox[t]);
                sm.raux[t] = aux[ibase + row];
#pragma unroll
                for (int k = 0; k < 4; ++k) sm.rt[t][k] = vertex_t(sm.rbox[t][2 * k], sm.rbox[t][2 * k + 1]);
            }
            if (col < m) {
                stage_oriented(boxes + (ibase + col) * 8, sm.cbox[t]);
                sm.caux[t] = aux[ibase + col];
#pragma unroll
                for (int k = 0; k < 4; ++k) sm.ct[t][k] = vertex_t(sm.cbox[t][2 * k], sm.cbox[t][2 * k + 1]);
            }
            sm.bits[t] = 0;
        }
        if (t == 0) {
            sm.qn = 0;
            sm.mqn = 0;
            sm.stat_pairs = 0;
        }
        __syncthreads();
        const int ncol = min(64, m - c0);
        // phase 1: 16 threads per row, 4 columns each; pre-filter against the alive columns, queue what needs the clip
        {
            constexpr int kCols = 64 / (kDiagThreads / (64 / kDiagSplit));  // columns per thread
            const int r = sub * (64 / kDiagSplit) + t / (64 / kCols), jq = (t % (64 / kCols)) * kCols;
            const unsigned lane = t & 31;
            const bool row_on = r0 + r < m && !((rdead >> r) & 1ull);
            const NmsAux P = sm.raux[row_on ? r : 0];
            const int jlo = max(jq, (cb == rb) ? r + 1 : 0), jhi = min(jq + kCols, ncol);
            unsigned npairs = 0;
            for (int u = 0; u < kCols; ++u) {
                const int j = jq + u;
                bool want = row_on && j >= jlo && j < jhi && !((cdead >> j) & 1ull);
                bool maybe = false;
                if (want) {
                    ++npairs;
                    const NmsAux& Q = sm.caux[j];
                    if ((P.area + Q.area) != 0.f) {
                        const int cls = pair_filter_quick(P, Q);
                        if (cls == 0) want = false;
                        if (cls == 2) {
                            want = false;
                            maybe = true;
                        }
                    }
                }
                queue_push(sm.queue, &sm.qn, want, static_cast<unsigned short>((r << 6) | j), lane);
                queue_push(sm.terms.tq, &sm.mqn, maybe, static_cast<unsigned short>((r << 6) | j), lane);
            }
            for (int o = 16; o > 0; o >>= 1) npairs += __shfl_xor_sync(0xffffffffu, npairs, o);
            if (lane == 0 && npairs) atomicAdd(&sm.stat_pairs, npairs);
        }
        __syncthreads();
        {
            // phase 1b: the pairs the cheap exits left undecided, one per lane (see nms_bcast_kernel)
            const int mq = sm.mqn;
            const unsigned lane = t & 31;
            for (int k0 = 0; k0 < mq; k0 += kDiagThreads) {
                const int k = k0 + t;
                bool want = k < mq;
                const unsigned short e = want ? sm.terms.tq[k] : static_cast<unsigned short>(0);
                if (want) {
                    const int r = e >> 6, j = e & 63;
                    if (pair_inter_is_zero(sm.raux[r], sm.caux[j], sm.rbox[r], sm.cbox[j])) want = false;
                }
                queue_push(sm.queue, &sm.qn, want, e, lane);
            }
            __syncthreads();
        }
        // phase 2: the queued pairs, term by term
        const int qn = sm.qn;
        clip_queue<false, kDiagThreads>(sm.queue, qn, sm.rbox, sm.cbox, sm.raux, sm.caux, sm.rt, sm.ct, sm.terms,
                                        sm.poly + t, thr, sm.bits, nullptr);
        if (t == 0) {
            atomicAdd(stats + kStDiagPairs, static_cast<u64>(sm.stat_pairs));
            atomicAdd(stats + kStDiagQueued, static_cast<u64>(qn));
        }
        if (t < 64) bits_out = sm.bits[t];
    }
    if (t >= sub * (64 / kDiagSplit) && t < (sub + 1) * (64 / kDiagSplit))  // this CTA's 16 rows of the block
        diag[(static_cast<size_t>(n) * kPanel + rb * 64 + t) * kPanelWords + cb] = bits_out;
    __threadfence();
    __syncthreads();
    if (t == 0) sm.last = (atomicAdd(&ctr[n], 1) == kDiagBlocks * kDiagSplit - 1);
    __syncthreads();
    const bool last = sm.last;
    __syncthreads();  // the tile is dead from here on: the sweep reuses its memory
    if (!last) return;
    __threadfence();

    // ---- last block of this image: sequential sweep inside the panel (64 rows at a time). All hit words of the panel
    // come to shared memory in one pass of independent loads first: the sweep is a chain of 8 dependent steps, and a
    // dependent L2 round trip per step and per later word (28 of them) was what it spent its 25 us on.
    __shared__ u64 s_rem[kPanelWords], s_kept[kPanelWords];
    u64* s_dg = reinterpret_cast<u64*>(smem_raw);
    const u64* dg = diag + static_cast<size_t>(n) * kPanel * kPanelWords;
    const int rows_in_panel = min(kPanel, m - base);
#pragma unroll 8
    for (int i = t; i < rows_in_panel * kPanelWords; i += kDiagThreads) s_dg[i] = __ldcg(dg + i);
    if (t < kPanelWords) {
        const int w = (base >> 6) + t;
        s_rem[t] = w < nblk ? rmv[w] : ~0ull;
        s_kept[t] = 0;
    }
    __syncthreads();
    for (int b = 0; b * 64 < rows_in_panel; ++b) {
        const int rows = min(64, rows_in_panel - b * 64);
        if (t == 0) {
            u64 cur = s_rem[b], alive = 0;
            if (rows < 64) cur |= ~0ull << rows;
            const u64* dw = s_dg + static_cast<size_t>(b) * 64 * kPanelWords + b;  // diagonal word of row r: dw[r * 8]
#pragma unroll 8
            for (int r = 0; r < 64; ++r) {
                if (!((cur >> r) & 1ull)) {
                    alive |= 1ull << r;
                    if (r < rows) cur |= dw[r * kPanelWords];
                }
            }
            s_kept[b] = alive;
        }
        __syncthreads();
        // survivors of this block suppress later blocks of the panel
        if (t < 64 && ((s_kept[b] >> t) & 1ull)) {
            for (int w = b + 1; w < kPanelWords; ++w) {
                const u64 v = s_dg[(static_cast<size_t>(b) * 64 + t) * kPanelWords + w];
                if (v) atomicOr(&s_rem[w], v);
            }
        }
        __syncthreads();
    }
    // publish: kept words, removed = not kept for the panel's own boxes, keep list in order
    int nk = nkeep[n];
    for (int w = 0; w < kPanelWords; ++w) {
        const u64 kw = s_kept[w];
        if (t < 64 && ((kw >> t) & 1ull)) keep[ibase + nk + __popcll(kw & ((1ull << t) - 1ull))] = base + w * 64 + t;
        nk += __popcll(kw);
    }
    if (t < kPanelWords) {
        pk[static_cast<size_t>(n) * kPanelWords + t] = s_kept[t];
        const int w = (base >> 6) + t;
        if (w < nblk) rmv[w] = ~s_kept[t];
    }
    if (t == 0) {
        nkeep[n] = nk;
        ctr[n] = 0;
    }
}

// ------------------------------------------------------------------------------------------------ broadcast
struct BcastSmem {
    float rbox[kRowChunk][8];
    NmsAux raux[kRowChunk];
    float cbox[kColChunk][8];
    NmsAux caux[kColChunk];
    float2 rctr[kRowChunk], cctr[kColChunk];  // 4 x centre (vertex sum) of the staged boxes: the "close pair" test
    float rt[kRowChunk][4], ct[kColChunk][4];  // t(v) of the staged boxes' vertices (term_is_zero)
    TermSmem terms;
    unsigned short queue[kRowChunk * kColChunk];
    unsigned char dead[kColChunk];
    unsigned int newdead[kColChunk];  // 0 -> 1 flags, touched with atomics while phase 2 runs
    int rows[kRowChunk];
    int qn;
    int mqn;  // pairs the cheap exits of the pre-filter left undecided (listed in terms.tq, which phase 2 reuses)
    unsigned stat_pairs;
    float2 poly[9 * kBcastThreads];  // tri_overlap's per-thread polygon columns
};

// Work item = (column chunk after the panel, row chunk of the panel's kept rows, image). How many there are is only
// known on the device (it depends on every image's box count and on how many rows of the panel survived), so the grid
// is a fixed number of resident CTAs that pull items from a global counter: a panel past the last box of every image
// costs one wave of CTAs that find zero items instead of tens of thousands of CTAs launched to exit (the capacity bound
// is 8 960 boxes per image, a typical image has a quarter of that), and the dynamic pull keeps the load balance the
// hardware's CTA dispatcher gave the one-CTA-per-item version.
__global__ void __launch_bounds__(kBcastThreads) nms_bcast_kernel(const float* __restrict__ boxes,
                                                              const NmsAux* __restrict__ aux,
                                                              const int* __restrict__ counts, int N, int max_sel,
                                                              int nblk, int panel, float thr, int pass,
                                                              float close_k,
                                                              const u64* __restrict__ pk, u64* __restrict__ removed,
                                                              u64* __restrict__ stats, int* __restrict__ work_ctr) {
    __shared__ BcastSmem sm;
    __shared__ int s_pref[kBcastImages + 1];  // exclusive prefix of the per-image item counts
    __shared__ int s_next;
    const int t = threadIdx.x;
    for (int n = t; n < N; n += kBcastThreads) {
        const int after = counts[n] - (panel + 1) * kPanel;
        int items = 0;
        if (after > 0) {
            int kept = 0;
#pragma unroll
            for (int w = 0; w < kPanelWords; ++w) kept += __popcll(pk[static_cast<size_t>(n) * kPanelWords + w]);
            items = ((after + kColChunk - 1) / kColChunk) * ((kept + kRowChunk - 1) / kRowChunk);
        }
        s_pref[n + 1] = items;
    }
    __syncthreads();
    if (t == 0) {
        s_pref[0] = 0;
        for (int n = 0; n < N; ++n) s_pref[n + 1] += s_pref[n];
    }
    __syncthreads();
    const int total_items = s_pref[N];
    int item = blockIdx.x;
    int n = 0;
    while (item < total_items) {
        while (item >= s_pref[n + 1]) ++n;  // items are handed out in increasing order
        const int m = counts[n];
        const int col_chunks = (m - (panel + 1) * kPanel + kColChunk - 1) / kColChunk;
        const int local = item - s_pref[n];
        const int c0 = (panel + 1) * kPanel + (local % col_chunks) * kColChunk;
        // the row chunk: kept rows number [k0, k0 + kRowChunk) of the panel
        const u64* pkn = pk + static_cast<size_t>(n) * kPanelWords;
        int total = 0;
        u64 kw[kPanelWords];
#pragma unroll
        for (int w = 0; w < kPanelWords; ++w) {
            kw[w] = pkn[w];
            total += __popcll(kw[w]);
        }
        const int k0 = (local / col_chunks) * kRowChunk;
        const int nrows = min(kRowChunk, total - k0);
        const size_t ibase = static_cast<size_t>(n) * max_sel;
        if (t < nrows) {
            // the (k0 + t)-th set bit of the 512-bit kept mask
            int want = k0 + t, w = 0;
            while (want >= __popcll(kw[w])) {
                want -= __popcll(kw[w]);
                ++w;
            }
            u64 v = kw[w];
            for (int i = 0; i < want; ++i) v &= v - 1;
            sm.rows[t] = panel * kPanel + w * 64 + (__ffsll(static_cast<long long>(v)) - 1);
        }
        if (t == 0) {
            sm.qn = 0;
            sm.mqn = 0;
            sm.stat_pairs = 0;
        }
        __syncthreads();
        if (t < nrows) {
            stage_oriented(boxes + (ibase + sm.rows[t]) * 8, sm.rbox[t]);
            sm.raux[t] = aux[ibase + sm.rows[t]];
            const float* b = sm.rbox[t];
            sm.rctr[t] = make_float2(b[0] + b[2] + b[4] + b[6], b[1] + b[3] + b[5] + b[7]);
#pragma unroll
            for (int k = 0; k < 4; ++k) sm.rt[t][k] = vertex_t(b[2 * k], b[2 * k + 1]);
        }
        u64* rmv = removed + static_cast<size_t>(n) * nblk;
        bool alive = false;
        if (t < kColChunk) {
            const int col = c0 + t;
            if (col < m) {
                alive = !((__ldcg(rmv + (col >> 6)) >> (col & 63)) & 1ull);
                if (alive) {
                    stage_oriented(boxes + (ibase + col) * 8, sm.cbox[t]);
                    sm.caux[t] = aux[ibase + col];
                    const float* b = sm.cbox[t];
                    sm.cctr[t] = make_float2(b[0] + b[2] + b[4] + b[6], b[1] + b[3] + b[5] + b[7]);
#pragma unroll
                    for (int k = 0; k < 4; ++k) sm.ct[t][k] = vertex_t(b[2 * k], b[2 * k + 1]);
                }
            }
            sm.dead[t] = alive ? 0 : 1;
            sm.newdead[t] = 0;
        }
        __syncthreads();
        // phase 1: kBcastThreads / kColChunk threads per column, a slice of the rows each
        {
            constexpr int kTpc = kBcastThreads / kColChunk, kRpt = (kRowChunk + kTpc - 1) / kTpc;
            const int j = t % kColChunk, part = t / kColChunk;
            const unsigned lane = t & 31;
            const bool col_on = !sm.dead[j];
            const NmsAux Q = sm.caux[col_on ? j : 0];
            const float2 qc = sm.cctr[col_on ? j : 0];
            unsigned npairs = 0;
            for (int u = 0; u < kRpt; ++u) {
                const int r = part * kRpt + u;
                bool want = col_on && r < nrows;
                bool maybe = false;
                if (want) {
                    const NmsAux& P = sm.raux[r];
                    // "close": squared centre distance below close_k x the smaller area (centres are kept x 4, hence
                    // the 16). A work-ordering heuristic only -- every pair is evaluated in exactly one of the passes
                    // unless its column has died in between.
                    const float dx = sm.rctr[r].x - qc.x, dy = sm.rctr[r].y - qc.y;
                    const bool close = dx * dx + dy * dy < 16.0f * close_k * fminf(P.area, Q.area);
                    if (close != (pass == 0)) {
                        want = false;
                    } else {
                        ++npairs;
                        // the cheap exits of the pre-filter here; what they leave undecided goes to a list of its own
                        // (pairs whose areas sum to 0 are always clipped: uni == 0 takes the reference's other formula)
                        if (pass != 0 && (P.area + Q.area) != 0.f) {
                            const int cls = pair_filter_quick(P, Q);
                            if (cls == 0) want = false;
                            if (cls == 2) {
                                want = false;
                                maybe = true;
                            }
                        }
                    }
                }
                queue_push(sm.queue, &sm.qn, want, static_cast<unsigned short>((r << 7) | j), lane);
                if (pass != 0) queue_push(sm.terms.tq, &sm.mqn, maybe, static_cast<unsigned short>((r << 7) | j), lane);
            }
            for (int o = 16; o > 0; o >>= 1) npairs += __shfl_xor_sync(0xffffffffu, npairs, o);
            if (lane == 0 && npairs) atomicAdd(&sm.stat_pairs, npairs);
        }
        __syncthreads();
        // phase 1b: the undecided pairs, one per lane -- every lane of a warp runs the expensive part of the pre-filter
        // (two divisions, up to 16 cross products) instead of the five that reached it inside the consult loop
        if (pass != 0) {
            const int mq = sm.mqn;
            const unsigned lane = t & 31;
            for (int k0 = 0; k0 < mq; k0 += kBcastThreads) {  // uniform trip count: every lane reaches queue_push
                const int k = k0 + t;
                bool want = k < mq;
                const unsigned short e = want ? sm.terms.tq[k] : static_cast<unsigned short>(0);
                if (want) {
                    const int r = e >> 7, j = e & 127;
                    if (pair_inter_is_zero(sm.raux[r], sm.caux[j], sm.rbox[r], sm.cbox[j])) want = false;
                }
                queue_push(sm.queue, &sm.qn, want, e, lane);
            }
            __syncthreads();
        }
        // phase 2: the queued pairs, term by term
        const int qn = sm.qn;
        clip_queue<true, kBcastThreads>(sm.queue, qn, sm.rbox, sm.cbox, sm.raux, sm.caux, sm.rt, sm.ct, sm.terms,
                                        sm.poly + t, thr, nullptr, sm.newdead);
        if (t == 0) {
            atomicAdd(stats + kStBcastPairs, static_cast<u64>(sm.stat_pairs));
            atomicAdd(stats + kStBcastQueued, static_cast<u64>(qn));
        }
        __syncthreads();
        if (t < kColChunk) {
            const bool newly = alive && sm.newdead[t];
            const unsigned bal = __ballot_sync(0xffffffffu, newly);
            if (bal && (t & 31) == 0) {
                const int cw = c0 + t;  // 32 columns of one warp share a 64-bit word
                atomicOr(rmv + (cw >> 6), static_cast<u64>(bal) << (cw & 63));
            }
        }
        // next item (everyone is done with the shared-memory tile once this barrier is passed)
        __syncthreads();
        if (t == 0) s_next = static_cast<int>(gridDim.x) + atomicAdd(work_ctr, 1);
        __syncthreads();
        item = s_next;
    }
}

// ------------------------------------------------------------------------------------------------ host
#define NMS_CHECK_LAUNCH(name)                                          \
    do {                                                                \
        cudaError_t e__ = cudaGetLastError();                           \
        if (e__ != cudaSuccess) {                                       \
            set_error("%s launch: %s", name, cudaGetErrorString(e__)); \
            return -1;                                                  \
        }                                                               \
    } while (0)

int run_nms(const float* nmsbox, const int* counts, int N, int max_sel, float thr, void* scratch, size_t scratch_bytes,
            int* keep, int* nkeep, cudaStream_t s, int64_t* launches) {
    const NmsLayout y = nms_layout(N, max_sel);
    if (y.total > scratch_bytes) {
        set_error("nms: scratch of %zu bytes is too small, need %zu", scratch_bytes, y.total);
        return -1;
    }
    uint8_t* b = static_cast<uint8_t*>(scratch);
    NmsAux* aux = reinterpret_cast<NmsAux*>(b + y.o_aux);
    u64* removed = reinterpret_cast<u64*>(b + y.o_removed);
    u64* diag = reinterpret_cast<u64*>(b + y.o_diag);
    u64* pk = reinterpret_cast<u64*>(b + y.o_pk);
    int* ctr = reinterpret_cast<int*>(b + y.o_ctr);
    u64* stats = reinterpret_cast<u64*>(b + y.o_stats);
    int* work = reinterpret_cast<int*>(b + y.o_work);
    const int groups = (N + kBcastImages - 1) / kBcastImages;
    // pass 0 of the broadcast takes the pairs with squared centre distance < close_k x min(area): tuning aid
    static float close_k = -1.f;
    if (close_k < 0.f) {
        const char* ev = getenv("DAFNE_NMS_CLOSE");
        close_k = ev ? static_cast<float>(atof(ev)) : 5.0f;
    }
    static DeviceOnce bcast_ctas_of;  // resident CTAs of the broadcast kernel, per device
    int dev = 0;
    int bcast_ctas = bcast_ctas_of.get(&dev);
    if (bcast_ctas == 0) {
        int sms = 0, per_sm = 0;
        cudaError_t q = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (q == cudaSuccess)
            q = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nms_bcast_kernel, kBcastThreads, 0);
        if (q != cudaSuccess || sms <= 0 || per_sm <= 0) {
            set_error("nms: occupancy query failed: %s", cudaGetErrorString(q));
            return -1;
        }
        bcast_ctas = sms * per_sm;
        bcast_ctas_of.set(dev, bcast_ctas);
    }
    // removed | pk | ctr | stats | work are contiguous: one clear
    cudaError_t e = cudaMemsetAsync(removed, 0, y.o_diag - y.o_removed, s);
    if (e == cudaSuccess) e = cudaMemsetAsync(nkeep, 0, static_cast<size_t>(N) * 4, s);
    if (e != cudaSuccess) {
        set_error("nms memset: %s", cudaGetErrorString(e));
        return -1;
    }
    nms_aux_kernel<<<dim3((max_sel + 127) / 128, N), 128, 0, s>>>(nmsbox, counts, max_sel, aux);
    NMS_CHECK_LAUNCH("nms_aux_kernel");
    const int panels = (max_sel + kPanel - 1) / kPanel;
    int nl = 1;
    for (int p = 0; p < panels; ++p) {
        nms_diag_kernel<<<dim3(kDiagBlocks * kDiagSplit, N), kDiagThreads, 0, s>>>(nmsbox, aux, counts, max_sel, y.nblk, p, thr, removed, diag,
                                                            pk, ctr, keep, nkeep, stats);
        NMS_CHECK_LAUNCH("nms_diag_kernel");
        ++nl;
        const int after = max_sel - (p + 1) * kPanel;
        for (int pass = 0; pass < 2; ++pass) {
            if (pass == 0 && !(close_k > 0.f)) continue;  // DAFNE_NMS_CLOSE=0: single pass (every pair is "far")
            for (int g = 0; after > 0 && g < groups; ++g) {
                const int n0 = g * kBcastImages, ng = N - n0 < kBcastImages ? N - n0 : kBcastImages;
                const long long most =
                    static_cast<long long>((after + kColChunk - 1) / kColChunk) * (kPanel / kRowChunk) * ng;
                const int grid = most < bcast_ctas ? static_cast<int>(most) : bcast_ctas;
                nms_bcast_kernel<<<grid, kBcastThreads, 0, s>>>(
                    nmsbox + static_cast<size_t>(n0) * max_sel * 8, aux + static_cast<size_t>(n0) * max_sel, counts + n0,
                    ng, max_sel, y.nblk, p, thr, pass, close_k, pk + static_cast<size_t>(n0) * kPanelWords,
                    removed + static_cast<size_t>(n0) * y.nblk, stats, work + (p * 2 + pass) * groups + g);
                NMS_CHECK_LAUNCH("nms_bcast_kernel");
                ++nl;
            }
        }
    }
    if (launches) *launches += nl;
    return 0;
}

int nms_read_stats(const void* scratch, int N, int max_sel, unsigned long long* host_out, cudaStream_t s) {
    const NmsLayout y = nms_layout(N, max_sel < 1 ? 1 : max_sel);
    cudaError_t e = cudaStreamSynchronize(s);
    if (e == cudaSuccess)
        e = cudaMemcpy(host_out, static_cast<const uint8_t*>(scratch) + y.o_stats, kNmsStats * 8, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) {
        set_error("nms_read_stats: %s", cudaGetErrorString(e));
        return -1;
    }
    return 0;
}

}  // namespace dafne
