// Polygon IoU of two quadrilaterals in the faithful fp32 arithmetic, plus the decision-preserving pre-filter the NMS
// kernels use. Device-only header; every translation unit that includes it MUST be compiled with
// -fmad=false -prec-div=true -prec-sqrt=true -ftz=false (see Makefile) so each operation is rounded once, in source
// order, exactly like oracle/polyiou_oracle.c.
//
//   algorithm ............ tools/prepare_dota/polyiou.cpp:8-133 (the only in-tree statement; double precision there)
//   call site ............ dafne/modeling/nms/nms.py:91 -> external poly_nms (fp32 transliteration of the above)
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace dafne {

struct P2 {
    float x, y;
};
// sig(d) = (d > eps) - (d < -eps) with the reference's DOUBLE eps = 1e-8 applied to a float: for a float d,
// (double)d > 1e-8  <=>  d > EPS_BELOW where EPS_BELOW is the largest float <= 1e-8 (there is no float in between).
__device__ __forceinline__ int sigf(float d) {
    const float eps_below = 9.99999993922529029e-09f;  // == (float)1e-8, which rounds down
    return (d > eps_below) - (d < -eps_below);
}
__device__ __forceinline__ bool same_pt(P2 a, P2 b) { return sigf(a.x - b.x) == 0 && sigf(a.y - b.y) == 0; }
__device__ __forceinline__ float cross3(P2 o, P2 a, P2 b) {
    return (a.x - o.x) * (b.y - o.y) - (b.x - o.x) * (a.y - o.y);
}
__device__ __forceinline__ float signed_area(P2* ps, int n) {
    float acc = 0.f;
    ps[n] = ps[0];
    for (int i = 0; i < n; i++) acc += ps[i].x * ps[i + 1].y - ps[i].y * ps[i + 1].x;
    return acc / 2.0f;
}
// ------------------------------------------------------------------------------------------------ polygon_cut
// polygon_cut (polyiou.cpp:58-71) restated for the GPU without changing one rounding:
//   * cross3(a, b, p[i]) is evaluated ONCE per vertex and reused as the `sj` of the previous edge and as the s1 / s2
//     of lineCross (polyiou.cpp:31-40) -- the reference recomputes the same expression on the same operands;
//   * sizes are static: a polygon of n vertices leaves polygon_cut with at most (#vertices kept) + (#sign changes)
//     <= n + n/2 points, so the three cuts of a triangle see at most 3 -> 4 -> 6 -> 9 vertices. Every loop is
//     unrolled over that bound with an `i < n` guard, the input polygon sits in registers, and the de-duplication
//     (polyiou.cpp:66-70) is streamed: each emitted point is compared with the point emitted before it and, if kept,
//     stored to the thread's private shared-memory column slots[k * stride] -- the only dynamically indexed access;
//   * the slot lineCross leaves unwritten when its denominator vanishes is (0, 0) (see oracle header).
// num / den, correctly rounded like the plain operator (den != 0 here). A ZERO numerator is the common case -- every
// crossing with an edge that starts or ends in the origin has the form (0 * s2 - v * 0) / (s2 - s1) -- and sends the
// hardware's division straight into its ~35-instruction slow path; IEEE gives it the value 0 with the sign
// sign(num) ^ sign(den) (den == den excludes a NaN denominator, for which the quotient is NaN).
__device__ __forceinline__ float div_exact(float num, float den) {
    if (num == 0.f && den == den)
        return __int_as_float((__float_as_int(num) ^ __float_as_int(den)) & static_cast<int>(0x80000000u));
    return num / den;
}

struct PolyEmit {
    float2* slots;
    int stride;
    int cnt, n;
    P2 prev, first, last;
    __device__ __forceinline__ void put(P2 e) {
        const bool keep = cnt == 0 || !same_pt(e, prev);  // tmp[i] vs tmp[i-1]
        if (keep) {
            slots[n * stride] = make_float2(e.x, e.y);
            if (n == 0) first = e;
            last = e;
            ++n;
        }
        prev = e;
        ++cnt;
    }
};

// Cuts the polygon p[0..n) (n <= NI, in registers) by the left half-plane of a->b; the result (<= NI + NI/2 points)
// is written to slots[0 .. return) in order.
template <int NI>
__device__ __forceinline__ int clip_left_static(const P2 (&p)[NI], int n, P2 a, P2 b, float2* slots, int stride) {
    float s[NI];
    int g[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        s[i] = cross3(a, b, p[i]);
        g[i] = sigf(s[i]);
    }
    PolyEmit em;
    em.slots = slots;
    em.stride = stride;
    em.cnt = 0;
    em.n = 0;
    em.prev = p[0];
    em.first = p[0];
    em.last = p[0];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        if (i < n) {
            // successor: i + 1, or vertex 0 after the last one (p[n] = p[0], polyiou.cpp:60)
            const bool wrap = (i + 1 >= NI) || (i + 1 >= n);
            const float s_nxt = wrap ? s[0] : s[(i + 1) % NI];
            const int g_nxt = wrap ? g[0] : g[(i + 1) % NI];
            if (g[i] > 0) em.put(p[i]);
            if (g[i] != g_nxt) {
                const P2 q = wrap ? p[0] : p[(i + 1) % NI];
                P2 x;
                x.x = 0.f;
                x.y = 0.f;
                const float den = s_nxt - s[i];
                if (sigf(den) != 0) {
                    x.x = div_exact(p[i].x * s_nxt - q.x * s[i], den);
                    x.y = div_exact(p[i].y * s_nxt - q.y * s[i], den);
                }
                em.put(x);
            }
        }
    }
    int m = em.n;
    // while (n > 1 && p[n-1] == p[0]) n--   (polyiou.cpp:70); the first round runs on registers
    if (m > 1 && same_pt(em.last, em.first)) {
        --m;
        while (m > 1) {
            const float2 l = slots[(m - 1) * stride];
            P2 lp;
            lp.x = l.x;
            lp.y = l.y;
            if (!same_pt(lp, em.first)) break;
            --m;
        }
    }
    return m;
}

template <int NI>
__device__ __forceinline__ void load_poly(P2 (&p)[NI], int n, const float2* slots, int stride) {
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        float2 v = make_float2(0.f, 0.f);
        if (i < n) v = slots[i * stride];
        p[i].x = v.x;
        p[i].y = v.y;
    }
}

// The same cut for a polygon of any size the algorithm can produce (<= 9 in, <= 13 out), looping over local arrays.
// Only the rare polygons that outgrow the static bounds chosen in tri_overlap come here.
static __device__ __noinline__ int clip_left_generic(float2* slots, int stride, int n, P2 a, P2 b) {
    P2 p[10], tmp[14];
    for (int i = 0; i < n; ++i) {
        const float2 v = slots[i * stride];
        p[i].x = v.x;
        p[i].y = v.y;
    }
    int m = 0;
    p[n] = p[0];
    const float s_first = n > 0 ? cross3(a, b, p[0]) : 0.f;
    float s_cur = s_first;
    int g_cur = sigf(s_cur);
    for (int i = 0; i < n; i++) {
        const float s_nxt = (i + 1 == n) ? s_first : cross3(a, b, p[i + 1]);
        const int g_nxt = sigf(s_nxt);
        if (g_cur > 0) tmp[m++] = p[i];
        if (g_cur != g_nxt) {
            P2 x;
            x.x = 0.f;
            x.y = 0.f;
            const float den = s_nxt - s_cur;
            if (sigf(den) != 0) {
                x.x = div_exact(p[i].x * s_nxt - p[i + 1].x * s_cur, den);
                x.y = div_exact(p[i].y * s_nxt - p[i + 1].y * s_cur, den);
            }
            tmp[m++] = x;
        }
        s_cur = s_nxt;
        g_cur = g_nxt;
    }
    n = 0;
    for (int i = 0; i < m; i++)
        if (i == 0 || !same_pt(tmp[i], tmp[i - 1])) p[n++] = tmp[i];
    while (n > 1 && same_pt(p[n - 1], p[0])) n--;
    for (int i = 0; i < n; ++i) slots[i * stride] = make_float2(p[i].x, p[i].y);
    return n;
}

// Signed overlap of triangles (O, a, b) and (O, c, d), O = origin (polyiou.cpp:74-89). `slots` is this thread's
// private column of >= 9 float2 in shared memory, `stride` elements apart.
//
// Static bounds: the first cut leaves [~O, v1, v2] (3 points) unless a vertex lies exactly on the ray, the second 3-4
// points, the third 3-5; those sizes get fully unrolled code (3 / 3 / 4 input vertices, 5 for the area), anything
// larger -- still the same arithmetic -- takes the looping fallback.
__device__ __forceinline__ float tri_overlap(P2 a, P2 b, P2 c, P2 d, float2* slots, int stride) {
    P2 o;
    o.x = 0.f;
    o.y = 0.f;
    const int s1 = sigf(cross3(o, a, b));
    const int s2 = sigf(cross3(o, c, d));
    if (s1 == 0 || s2 == 0) return 0.f;
    if (s1 == -1) {
        P2 t = a;
        a = b;
        b = t;
    }
    if (s2 == -1) {
        P2 t = c;
        c = d;
        d = t;
    }
    // Exact shortcut (same result bits as the generic path below): if a and b are both strictly right of the ray
    // O->c, the first clip keeps no vertex and both intersection points it appends are exactly (0,0)
    // ((0*s2 - v*0)/(s2 - 0)); every later clip then sees a single zero point and the area is exactly 0.
    const float sa = c.x * a.y - a.x * c.y;  // == cross3(O, c, a): x - 0 is exact
    const float sb = c.x * b.y - b.x * c.y;
    const int ga = sigf(sa), gb = sigf(sb);
    if (ga < 0 && gb < 0) return 0.f;
    int n;
    // First cut (left of O -> c) of [O, a, b]. O lies on the line (sign class 0), so whenever a and b are off the
    // line the cut is one of three fixed shapes, and the intersection points with an edge through O are zero points
    // ((0 * s2 - v * 0) / (s2 - 0)); the sign of such a zero never reaches the result (it only meets +-0 products,
    // x - c with c != 0, and same_pt / sig, which treat +-0 alike):
    //   a, b left:   tmp = [z, a, b, z']    -> [z, a, b]       a left, b right:  tmp = [z, a, x, z'] -> [z, a, x]
    //   a right, b left:  tmp = [z, x, b, z'] -> [z, x, b]      (x = lineCross(a, b), z' dropped by the closing check)
    // provided the de-duplication (polyiou.cpp:66-70) finds the three points distinct -- checked; anything else
    // (a vertex on the line, coincident points) takes the generic cut.
    bool fast = false;
    // (non-finite coordinates would turn v * 0 into NaN: generic cut)
    if (ga != 0 && gb != 0 && fabsf(a.x) + fabsf(a.y) + fabsf(b.x) + fabsf(b.y) < INFINITY) {
        P2 z, p1 = a, p2 = b;
        z.x = 0.f;
        z.y = 0.f;
        if (ga != gb) {
            P2 x;
            x.x = 0.f;
            x.y = 0.f;
            const float den = sb - sa;
            if (sigf(den) != 0) {
                x.x = div_exact(a.x * sb - b.x * sa, den);
                x.y = div_exact(a.y * sb - b.y * sa, den);
            }
            if (ga > 0)
                p2 = x;
            else
                p1 = x;
        }
        if (!same_pt(p1, z) && !same_pt(p2, p1) && !same_pt(p2, z)) {
            slots[0] = make_float2(0.f, 0.f);
            slots[stride] = make_float2(p1.x, p1.y);
            slots[2 * stride] = make_float2(p2.x, p2.y);
            n = 3;
            fast = true;
        }
    }
    if (!fast) {
        P2 p3[3];
        p3[0] = o;
        p3[1] = a;
        p3[2] = b;
        n = clip_left_static<3>(p3, 3, o, c, slots, stride);
    }
    if (n <= 3) {
        P2 p3[3];
        load_poly<3>(p3, n, slots, stride);
        n = clip_left_static<3>(p3, n, c, d, slots, stride);
    } else {
        n = clip_left_generic(slots, stride, n, c, d);
    }
    if (n <= 4) {
        P2 p4[4];
        load_poly<4>(p4, n, slots, stride);
        n = clip_left_static<4>(p4, n, d, o, slots, stride);
    } else {
        n = clip_left_generic(slots, stride, n, d, o);
    }
    // shoelace over the closed polygon (polyiou.cpp:23-30)
    float acc = 0.f;
    if (n <= 5) {
        P2 q[5];
        load_poly<5>(q, n, slots, stride);
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            if (i < n) {
                const bool wrap = (i + 1 >= 5) || (i + 1 >= n);
                const P2 nx = wrap ? q[0] : q[(i + 1) % 5];
                acc += q[i].x * nx.y - q[i].y * nx.x;
            }
        }
    } else {
        const float2 q0 = slots[0];
        float2 cur = q0;
        for (int i = 0; i < n; ++i) {
            const float2 nx = (i + 1 == n) ? q0 : slots[(i + 1) * stride];
            acc += cur.x * nx.y - cur.y * nx.x;
            cur = nx;
        }
    }
    float res = fabsf(acc / 2.0f);
    if (s1 * s2 == -1) res = -res;
    return res;
}
__device__ __forceinline__ void load_oriented(const float* pa, P2* p) {
    for (int i = 0; i < 4; i++) {
        p[i].x = pa[2 * i];
        p[i].y = pa[2 * i + 1];
    }
    if (signed_area(p, 4) < 0.f) {
        P2 t = p[0];
        p[0] = p[3];
        p[3] = t;
        t = p[1];
        p[1] = p[2];
        p[2] = t;
    }
    p[4] = p[0];
}
static __device__ __noinline__ float iou_poly_f32(const float* pa, const float* qa, float2* slots, int stride) {
    P2 p[6], q[6];
    load_oriented(pa, p);
    load_oriented(qa, q);
    float inter = 0.f;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) inter += tri_overlap(p[i], p[i + 1], q[j], q[j + 1], slots, stride);
    const float a1 = fabsf(signed_area(p, 4));
    const float a2 = fabsf(signed_area(q, 4));
    const float uni = a1 + a2 - inter;
    if (uni == 0.f) return (inter + 1.f) / (uni + 1.f);
    return inter / uni;
}

// ------------------------------------------------------------------------------------------------ NMS pre-filter
// Per-box scalars from which `pair_inter_is_zero(P, Q)` PROVES, for most geometrically separated pairs, that the
// fp32 arithmetic above yields inter == 0 exactly (hence IoU == 0 <= thr unless both areas are 0) without running
// it. This is not a bounding-box reject: with class-shifted coordinates the fp32 IoU of disjoint boxes is in general
// NOT zero; the filter only fires where every one of the 16 triangle pairs provably collapses.
//
// Notation for a point v with v.x, v.y >= 1:  s(v) = v.x + v.y,  t(v) = (v.y - v.x) / s(v)  (monotone in the polar
// angle), and for any two such points  c.x*v.y - v.x*c.y = 0.5 * s(c) * s(v) * (t(v) - t(c))  exactly.
//
// Case A, P clockwise of Q (thi(P) + m < tlo(Q)): for every vertex v of P and c of Q the rounded value
// fl(fl(c.x v.y) - fl(v.x c.y)) <= -(0.5 m' - u)(1 - u) s(c) s(v) < -1e-8 (u = 2^-24, m' = m minus the rounding
// slack folded into tlo/thi), so in every tri_overlap both P vertices are strictly right of O->c: the shortcut
// above returns exactly 0 sixteen times.
//
// Case A', P counter-clockwise of Q (tlo(P) - thi(Q) = gap > 0): the first clip is the identity on [O, a, b]; the
// second (line c->d) yields a polygon whose vertices are O-like exact zero points and points of the form a, b,
// k*a, k*b (k = sO / (sO - s_v) in (0, 1]) or an affine combination a + mu (b - a), mu in (-1, 2), each with
// relative rounding error <= 2^-22; the third clip (line d->O) evaluates cross3(d, O, x) for them. The inequality
// tested below makes every one of those rounded values < -1e-8 (derivation in DESIGN.md section "NMS pre-filter"),
// so each clip-3 intersection is (0*s2 - x*0)/(s2 - 0) = 0, the polygon collapses to zero points and the area is
// exactly 0. kq lower-bounds k through the distance of the origin from Q's edge lines; ext widens P's t-interval by
// one edge length for the affine combinations.
struct NmsAux {
    float tlo, thi;    // t-interval of the vertices, widened by the rounding slack; (-inf, +inf) if not eligible
    float smin, smax;  // range of s over the vertices (widened)
    float ext;         // 2 * D / (smin - D), D = longest edge in L1; +inf if not eligible
    float sminx;       // smin - D
    float kq;          // min over edges of cross3(c, d, O) / Linf(d - c); 0 if not eligible as the clipping box
    float area;        // |shoelace| of the (re)oriented quad, the algorithm's own a1 / a2
};

__device__ __forceinline__ NmsAux nms_aux_of(const float* box) {
    NmsAux a;
    P2 p[6];
    load_oriented(box, p);
    a.area = fabsf(signed_area(p, 4));
    bool ok = true;
    float tlo = INFINITY, thi = -INFINITY, smin = INFINITY, smax = 0.f, dmax = 0.f, kq = INFINITY;
    for (int i = 0; i < 4; ++i) {
        const float x = p[i].x, y = p[i].y;
        ok = ok && (x >= 1.0f) && (y >= 1.0f) && (x <= 1.0e7f) && (y <= 1.0e7f);  // also false for NaN
        const float s = x + y;
        const float t = (y - x) / s;
        tlo = fminf(tlo, t);
        thi = fmaxf(thi, t);
        smin = fminf(smin, s);
        smax = fmaxf(smax, s);
        const P2 u = p[i], w = p[i + 1];
        dmax = fmaxf(dmax, fabsf(w.x - u.x) + fabsf(w.y - u.y));
        // this edge as a clipping edge of Q: orient (c, d) counter-clockwise about the origin like tri_overlap
        const int s2 = sigf(u.x * w.y - w.x * u.y);
        if (s2 != 0) {
            const P2 c = s2 > 0 ? u : w, d = s2 > 0 ? w : u;
            const float sO = (d.x - c.x) * (0.f - c.y) - (0.f - c.x) * (d.y - c.y);  // cross3(c, d, O)
            const float linf = fmaxf(fabsf(d.x - c.x), fabsf(d.y - c.y));
            if (!(sO > 1.0f) || !(linf > 0.f))
                kq = 0.f;
            else
                kq = fminf(kq, sO / linf);
        }
    }
    if (!ok) {
        a.tlo = -INFINITY;
        a.thi = INFINITY;
        a.smin = 1.f;
        a.smax = INFINITY;
        a.ext = INFINITY;
        a.sminx = 1.f;
        a.kq = 0.f;
        return a;
    }
    a.tlo = tlo - 1.0e-6f;
    a.thi = thi + 1.0e-6f;
    a.smin = smin * 0.999999f;
    a.smax = smax * 1.000001f;
    const float D = dmax * 1.000001f;
    if (a.smin > 2.0f * D) {
        a.sminx = a.smin - D;
        a.ext = 2.0f * D / a.sminx * 1.000001f;
    } else {
        a.sminx = 1.f;
        a.ext = INFINITY;
    }
    a.kq = (kq == INFINITY) ? 0.f : kq * 0.999999f;
    return a;
}

// true => the faithful arithmetic gives inter == 0 exactly for IoU(P = higher-scored box, Q = lower-scored box).
// pbox / qbox: the boxes (re)oriented like iou_poly_f32 does (load_oriented), 8 floats each.
//
// Case A'' (tight form of A'): the margin `ext` of case A' only pays for intersection points that EXTRAPOLATE an edge
// of P, and lineCross extrapolates only when it is called for a vertex whose cross3(c, d, v) lies within eps of 0 (sign
// class 0 next to a sign class +-1 of the same numeric sign). If no vertex of P is within eps of any oriented edge
// line of Q -- checked below with the very expression polygon_cut evaluates -- every intersection point of the second
// cut is a convex combination of two consecutive vertices (weights s2/(s2-s1), -s1/(s2-s1) in (0,1), no cancellation
// in numerator or denominator, relative error <= 4.5 u), stays inside P's t-interval, and the inequality of case A'
// holds with ext = 0.
__device__ __forceinline__ bool pair_inter_is_zero(const NmsAux& P, const NmsAux& Q, const float* pbox,
                                                   const float* qbox) {
    if (P.thi + 1.0e-5f < Q.tlo) return true;  // case A
    const float gap0 = (P.tlo - Q.thi) - 4.0e-6f;
    if (!(gap0 > 0.f)) return false;
    const float kmin = Q.kq / (Q.kq + 1.002f * (P.smax + Q.smax));
    const float rho = Q.smax / P.sminx;
    const float rhs = 2.4e-7f * (4.0f * P.smax / P.sminx + rho + 2.0f) + 5.0e-9f;
    const float gapx = gap0 - P.ext;
    if (gapx > 0.f && kmin * (0.5f * gapx - 2.4e-7f) * 0.999f > rhs) return true;  // case A'
    if (!(P.ext < INFINITY)) return false;                                           // sminx is not a bound
    if (!(kmin * (0.5f * gap0 - 2.4e-7f) * 0.999f > rhs)) return false;
    // case A'': no vertex of P within eps of an oriented edge line of Q
    P2 o;
    o.x = 0.f;
    o.y = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        P2 c, d;
        c.x = qbox[2 * j];
        c.y = qbox[2 * j + 1];
        d.x = qbox[2 * ((j + 1) & 3)];
        d.y = qbox[2 * ((j + 1) & 3) + 1];
        const int s2 = sigf(cross3(o, c, d));
        if (s2 == 0) continue;  // tri_overlap returns 0 for this edge before any cut
        if (s2 == -1) {
            const P2 t = c;
            c = d;
            d = t;
        }
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            P2 pv;
            pv.x = pbox[2 * v];
            pv.y = pbox[2 * v + 1];
            if (sigf(cross3(c, d, pv)) == 0) return false;
        }
    }
    return true;
}

// The two cheap exits of pair_inter_is_zero, for callers that evaluate the rest on a compacted list (the NMS kernels: in
// the consult loops only a few lanes of a warp get past these two tests, and the divisions and the 16 cross products
// that follow ran at 5 of 32 lanes): 0 = inter is exactly 0 (case A), 1 = the pair has to be clipped (no gap), 2 =
// undecided -- pair_inter_is_zero decides it.
__device__ __forceinline__ int pair_filter_quick(const NmsAux& P, const NmsAux& Q) {
    if (P.thi + 1.0e-5f < Q.tlo) return 0;
    const float gap0 = (P.tlo - Q.thi) - 4.0e-6f;
    if (!(gap0 > 0.f)) return 1;
    return 2;
}

// ------------------------------------------------------------------------------------------------ per-term filter
// The same proofs, one term at a time. A pair that reaches the clip has, on average, 8 of its 16 signed triangle
// overlaps equal to zero (measured on class-shifted boxes: 4 where the edge triangle of P lies clockwise of Q's --
// tri_overlap's own first shortcut -- and 4 where it lies counter-clockwise of it); `term_is_zero` decides that
// BEFORE a lane is spent on the term, so the NMS kernels only queue the terms that have to be clipped.
//
//   a, b, c, d: the term's vertices as iou_poly_f32 passes them to tri_overlap (edge i of P, edge j of Q, both
//   polygons already oriented by load_oriented); ta .. td: t(v) = (v.y - v.x) / (v.x + v.y) of those vertices, computed
//   with exactly this expression. Returns true only if tri_overlap(a, b, c, d) is exactly (+-)0:
//     * s1 == 0 or s2 == 0: tri_overlap returns 0 before any cut (polyiou.cpp:77-79);
//     * both P vertices strictly right of O->c: tri_overlap's shortcut (exact by construction);
//     * case A'' of pair_inter_is_zero for THIS term: the proof never looks at another term -- the term's own gap
//       min(t(a), t(b)) - max(t(c), t(d)) (each t widened by 1e-6 like NmsAux does) replaces the boxes' gap, the bounds
//       smax / sminx / kq of the boxes cover the term's vertices and its edge, and "no vertex of P within eps of an
//       oriented edge line of Q" is needed for the term's own two vertices and its own edge only.
struct TermPairCtx {
    float kmin, rhs;  // of the pair; kmin == 0 disables the counter-clockwise case
};
__device__ __forceinline__ TermPairCtx term_pair_ctx(const NmsAux& P, const NmsAux& Q) {
    TermPairCtx x;
    x.kmin = 0.f;
    x.rhs = 1.f;
    if (P.ext < INFINITY && Q.kq > 0.f) {  // both boxes eligible (coordinates in [1, 1e7]), sminx is a valid bound
        x.kmin = Q.kq / (Q.kq + 1.002f * (P.smax + Q.smax));
        const float rho = Q.smax / P.sminx;
        x.rhs = 2.4e-7f * (4.0f * P.smax / P.sminx + rho + 2.0f) + 5.0e-9f;
    }
    return x;
}
__device__ __forceinline__ bool term_is_zero(const TermPairCtx& x, P2 a, P2 b, P2 c, P2 d, float ta, float tb, float tc,
                                             float td) {
    const int s1 = sigf(a.x * b.y - b.x * a.y);  // == cross3(O, a, b): v - 0 is exact
    const int s2 = sigf(c.x * d.y - d.x * c.y);
    if (s1 == 0 || s2 == 0) return true;
    if (s1 == -1) {
        const P2 t = a;
        a = b;
        b = t;
    }
    if (s2 == -1) {
        const P2 t = c;
        c = d;
        d = t;
    }
    const int ga = sigf(c.x * a.y - a.x * c.y), gb = sigf(c.x * b.y - b.x * c.y);
    if (ga < 0 && gb < 0) return true;
    if (!(x.kmin > 0.f)) return false;
    const float gap0 = ((fminf(ta, tb) - 1.0e-6f) - (fmaxf(tc, td) + 1.0e-6f)) - 4.0e-6f;
    if (!(gap0 > 0.f)) return false;
    if (!(x.kmin * (0.5f * gap0 - 2.4e-7f) * 0.999f > x.rhs)) return false;
    return sigf(cross3(c, d, a)) != 0 && sigf(cross3(c, d, b)) != 0;
}
__device__ __forceinline__ float vertex_t(float vx, float vy) { return (vy - vx) / (vx + vy); }

}  // namespace dafne
