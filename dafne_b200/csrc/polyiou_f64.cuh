// Polygon IoU of two quadrilaterals in DOUBLE precision: the arithmetic of the reference's SWIG module `polyiou`
// (tools/prepare_dota/polyiou.cpp:8-133), which its patch-merge step calls on the host
// (dafne/utils/ResultMerge_multi_process.py:61-122, py_cpu_nms_poly_fast). Operation for operation like
// oracle/polyiou_oracle.c compiled with REAL = double, which is pinned against the reference's own polyiou.cpp
// (oracle/_ref). Device-only header; the including translation unit MUST be compiled with -fmad=false
// -prec-div=true (see Makefile) so that every product, difference and quotient is rounded once, in source order.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace dafne {
namespace f64 {

struct P2d {
    double x, y;
};
__device__ __forceinline__ int sig(double d) { return (d > 1e-8) - (d < -1e-8); }  // polyiou.cpp:9-12
__device__ __forceinline__ bool same_pt(P2d a, P2d b) { return sig(a.x - b.x) == 0 && sig(a.y - b.y) == 0; }
__device__ __forceinline__ double cross3(P2d o, P2d a, P2d b) {  // polyiou.cpp:20-22
    return (a.x - o.x) * (b.y - o.y) - (b.x - o.x) * (a.y - o.y);
}
__device__ __forceinline__ double signed_area(P2d* ps, int n) {  // polyiou.cpp:23-30
    double acc = 0.0;
    ps[n] = ps[0];
    for (int i = 0; i < n; i++) acc += ps[i].x * ps[i + 1].y - ps[i].y * ps[i + 1].x;
    return acc / 2.0;
}
// polygon_cut (polyiou.cpp:58-71) with lineCross (polyiou.cpp:31-40) inlined; cross3(a, b, p[i]) is evaluated once
// per vertex and reused where the reference recomputes the same expression on the same operands. The slot lineCross
// leaves unwritten when its denominator vanishes is (0, 0) (see oracle header).
__device__ __forceinline__ void clip_left(P2d* p, int* n_io, P2d a, P2d b) {
    P2d tmp[16];
    int n = *n_io, m = 0;
    p[n] = p[0];
    const double s_first = n > 0 ? cross3(a, b, p[0]) : 0.0;
    double s_cur = s_first;
    int g_cur = sig(s_cur);
    for (int i = 0; i < n; i++) {
        const double s_nxt = (i + 1 == n) ? s_first : cross3(a, b, p[i + 1]);
        const int g_nxt = sig(s_nxt);
        if (g_cur > 0) tmp[m++] = p[i];
        if (g_cur != g_nxt) {
            P2d x;
            x.x = 0.0;
            x.y = 0.0;
            const double den = s_nxt - s_cur;
            if (sig(den) != 0) {
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
// signed overlap of triangles (O, a, b) and (O, c, d), O = origin (polyiou.cpp:74-89); a cut of an n-gon leaves at
// most n + n/2 points, so 3 -> 4 -> 6 -> 9 (+1 closing slot)
__device__ __forceinline__ double tri_overlap(P2d a, P2d b, P2d c, P2d d) {
    P2d o;
    o.x = 0.0;
    o.y = 0.0;
    const int s1 = sig(cross3(o, a, b));
    const int s2 = sig(cross3(o, c, d));
    if (s1 == 0 || s2 == 0) return 0.0;
    if (s1 == -1) {
        const P2d t = a;
        a = b;
        b = t;
    }
    if (s2 == -1) {
        const P2d t = c;
        c = d;
        d = t;
    }
    P2d p[12];
    int n = 3;
    p[0] = o;
    p[1] = a;
    p[2] = b;
    clip_left(p, &n, o, c);
    clip_left(p, &n, c, d);
    clip_left(p, &n, d, o);
    double res = fabs(signed_area(p, n));
    if (s1 * s2 == -1) res = -res;
    return res;
}
// iou_poly (polyiou.cpp:91-133): both quads re-oriented counter-clockwise, 16 signed triangle overlaps,
// union == 0 -> (inter + 1) / (union + 1)
__device__ __noinline__ double iou_poly(const double* pa, const double* qa) {
    P2d p[6], q[6];
    for (int i = 0; i < 4; i++) {
        p[i].x = pa[2 * i];
        p[i].y = pa[2 * i + 1];
        q[i].x = qa[2 * i];
        q[i].y = qa[2 * i + 1];
    }
    if (signed_area(p, 4) < 0.0) {
        P2d t = p[0];
        p[0] = p[3];
        p[3] = t;
        t = p[1];
        p[1] = p[2];
        p[2] = t;
    }
    if (signed_area(q, 4) < 0.0) {
        P2d t = q[0];
        q[0] = q[3];
        q[3] = t;
        t = q[1];
        q[1] = q[2];
        q[2] = t;
    }
    p[4] = p[0];
    q[4] = q[0];
    double inter = 0.0;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) inter += tri_overlap(p[i], p[i + 1], q[j], q[j + 1]);
    const double a1 = fabs(signed_area(p, 4));
    const double a2 = fabs(signed_area(q, 4));
    const double uni = a1 + a2 - inter;
    if (uni == 0.0) return (inter + 1.0) / (uni + 1.0);
    return inter / uni;
}

}  // namespace f64
}  // namespace dafne
