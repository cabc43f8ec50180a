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
__device__ __forceinline__ int line_cross(P2 a, P2 b, P2 c, P2 d, P2* out) {
    const float s1 = cross3(a, b, c);
    const float s2 = cross3(a, b, d);
    if (sigf(s1) == 0 && sigf(s2) == 0) return 2;
    if (sigf(s2 - s1) == 0) return 0;
    out->x = (c.x * s2 - d.x * s1) / (s2 - s1);
    out->y = (c.y * s2 - d.y * s1) / (s2 - s1);
    return 1;
}
// polygon_cut (polyiou.cpp:58-71). cross3(a, b, p[i]) is evaluated ONCE per vertex and reused as the `sj` of the
// previous edge and as the s1 / s2 of lineCross (polyiou.cpp:31-40): the reference recomputes the same expression on
// the same operands, so the values are identical and no rounding changes.
__device__ __forceinline__ void clip_left(P2* p, int* n_io, P2 a, P2 b) {
    P2 tmp[24];
    int n = *n_io, m = 0;
    p[n] = p[0];
    const float s_first = cross3(a, b, p[0]);
    float s_cur = s_first;
    int g_cur = sigf(s_cur);
    for (int i = 0; i < n; i++) {
        const float s_nxt = (i + 1 == n) ? s_first : cross3(a, b, p[i + 1]);
        const int g_nxt = sigf(s_nxt);
        if (g_cur > 0) tmp[m++] = p[i];
        if (g_cur != g_nxt) {
            // lineCross(a, b, p[i], p[i+1]): both signs zero cannot happen here; an unwritten slot is (0, 0)
            P2 x;
            x.x = 0.f;
            x.y = 0.f;
            const float den = s_nxt - s_cur;
            if (sigf(den) != 0) {
                x.x = (p[i].x * s_nxt - p[i + 1].x * s_cur) / den;
                x.y = (p[i].y * s_nxt - p[i + 1].y * s_cur) / den;
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
    *n_io = n;
}
__device__ __forceinline__ float tri_overlap(P2 a, P2 b, P2 c, P2 d) {
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
    if (sigf(c.x * a.y - a.x * c.y) < 0 && sigf(c.x * b.y - b.x * c.y) < 0) return 0.f;
    P2 p[12];
    int n = 3;
    p[0] = o;
    p[1] = a;
    p[2] = b;
    clip_left(p, &n, o, c);
    clip_left(p, &n, c, d);
    clip_left(p, &n, d, o);
    float res = fabsf(signed_area(p, n));
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
static __device__ __noinline__ float iou_poly_f32(const float* pa, const float* qa) {
    P2 p[6], q[6];
    load_oriented(pa, p);
    load_oriented(qa, q);
    float inter = 0.f;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) inter += tri_overlap(p[i], p[i + 1], q[j], q[j + 1]);
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

// true => the faithful arithmetic gives inter == 0 exactly for IoU(P = higher-scored box, Q = lower-scored box)
__device__ __forceinline__ bool pair_inter_is_zero(const NmsAux& P, const NmsAux& Q) {
    if (P.thi + 1.0e-5f < Q.tlo) return true;  // case A
    const float gapx = (P.tlo - Q.thi) - 4.0e-6f - P.ext;
    if (gapx > 0.f) {  // case A'
        const float kmin = Q.kq / (Q.kq + 1.002f * (P.smax + Q.smax));
        const float rho = Q.smax / P.sminx;
        const float lhs = kmin * (0.5f * gapx - 2.4e-7f) * 0.999f;
        const float rhs = 2.4e-7f * (4.0f * P.smax / P.sminx + rho + 2.0f) + 5.0e-9f;
        return lhs > rhs;
    }
    return false;
}

// Decision IoU(P, Q) > thr with the pre-filter in front (bit-identical decisions to iou_poly_f32(...) > thr).
__device__ __forceinline__ bool suppresses(const float* pbox, const NmsAux& P, const float* qbox, const NmsAux& Q,
                                           float thr) {
    if (pair_inter_is_zero(P, Q) && (P.area + Q.area) != 0.f) return false;  // IoU = 0 / (a1 + a2) = 0 <= thr
    return iou_poly_f32(pbox, qbox) > thr;
}

}  // namespace dafne
