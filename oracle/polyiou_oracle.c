/*
 * ORACLE (test infrastructure only -- never linked, imported or executed by the product path).
 *
 * CPU restatement of the polygon-IoU arithmetic and of the greedy rotated NMS the reference runs on every image.
 *   algorithm source ......... tools/prepare_dota/polyiou.cpp:8-133 (double precision; the only in-tree statement)
 *   call site ................ dafne/modeling/nms/nms.py:91  poly_gpu_nms(dets[n,9], thresh, device_id)
 *   third-party dependency ... poly_nms from CAPTAIN-WHU/DOTA_devkit @ 99388551054be9a6dabb01c8bb2a7eb562d57b4f
 *                              (Dockerfile:37-43); NOT vendored under the reference. Its published algorithm is a
 *                              float transliteration of polyiou.cpp (float points, the same double eps = 1e-8 and the
 *                              same (inter+1)/(union+1) degenerate branch), a 64x64-tiled pairwise `IoU > thresh`
 *                              bitmask over boxes pre-sorted by descending score, and a serial host sweep.
 *
 * The whole file is compiled twice (REAL = float and REAL = double) with -ffp-contract=off so that every product,
 * difference and quotient is rounded exactly once, in source order; the CUDA kernel is compiled with -fmad=false and
 * must agree with the float instantiation bit for bit. The double instantiation is pinned against the reference's
 * own polyiou.cpp compiled into oracle/_ref/ (see oracle/Makefile, tests/test_oracle_polyiou.py).
 *
 * One behaviour is DEFINED here because the reference leaves it undefined: polygon_cut() appends the output slot of
 * lineCross() even when lineCross() returns without writing it (polyiou.cpp:63 with :36); that slot is taken to be
 * (0, 0).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifndef REAL
#define REAL float
#endif
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

typedef struct {
    REAL x, y;
} FN(pt);
#define PT FN(pt)

static const double FN(k_eps) = 1e-8;

static int FN(sig)(REAL d) { return ((double)d > FN(k_eps)) - ((double)d < -FN(k_eps)); } /* polyiou.cpp:9-12 */

static int FN(same_pt)(PT a, PT b) { /* polyiou.cpp:16-18 */
    return FN(sig)(a.x - b.x) == 0 && FN(sig)(a.y - b.y) == 0;
}

static REAL FN(cross3)(PT o, PT a, PT b) { /* polyiou.cpp:20-22 */
    return (a.x - o.x) * (b.y - o.y) - (b.x - o.x) * (a.y - o.y);
}

/* shoelace; writes the closing vertex like the reference does (polyiou.cpp:23-30) */
static REAL FN(signed_area)(PT* ps, int n) {
    REAL acc = 0;
    int i;
    ps[n] = ps[0];
    for (i = 0; i < n; i++) acc += ps[i].x * ps[i + 1].y - ps[i].y * ps[i + 1].x;
    return (REAL)(acc / 2.0);
}

/* intersection of line ab with line cd (polyiou.cpp:31-40); *out untouched unless the result is 1 */
static int FN(line_cross)(PT a, PT b, PT c, PT d, PT* out) {
    REAL s1 = FN(cross3)(a, b, c);
    REAL s2 = FN(cross3)(a, b, d);
    if (FN(sig)(s1) == 0 && FN(sig)(s2) == 0) return 2;
    if (FN(sig)(s2 - s1) == 0) return 0;
    out->x = (c.x * s2 - d.x * s1) / (s2 - s1);
    out->y = (c.y * s2 - d.y * s1) / (s2 - s1);
    return 1;
}

/* keep the part of polygon p (n vertices) to the left of a->b, in place (polyiou.cpp:58-71) */
static void FN(clip_left)(PT* p, int* n_io, PT a, PT b) {
    PT tmp[24];
    int n = *n_io, m = 0, i;
    p[n] = p[0];
    for (i = 0; i < n; i++) {
        int si = FN(sig)(FN(cross3)(a, b, p[i]));
        int sj = FN(sig)(FN(cross3)(a, b, p[i + 1]));
        if (si > 0) tmp[m++] = p[i];
        if (si != sj) {
            tmp[m].x = 0; /* DEFINED: see file header */
            tmp[m].y = 0;
            FN(line_cross)(a, b, p[i], p[i + 1], &tmp[m]);
            m++;
        }
    }
    n = 0;
    for (i = 0; i < m; i++)
        if (i == 0 || !FN(same_pt)(tmp[i], tmp[i - 1])) p[n++] = tmp[i];
    while (n > 1 && FN(same_pt)(p[n - 1], p[0])) n--;
    *n_io = n;
}

/* signed overlap of triangles (O,a,b) and (O,c,d), O = origin (polyiou.cpp:74-89) */
static REAL FN(tri_overlap)(PT a, PT b, PT c, PT d) {
    PT o, p[12], t;
    int n = 3, s1, s2;
    REAL res;
    o.x = 0;
    o.y = 0;
    s1 = FN(sig)(FN(cross3)(o, a, b));
    s2 = FN(sig)(FN(cross3)(o, c, d));
    if (s1 == 0 || s2 == 0) return 0;
    if (s1 == -1) {
        t = a;
        a = b;
        b = t;
    }
    if (s2 == -1) {
        t = c;
        c = d;
        d = t;
    }
    p[0] = o;
    p[1] = a;
    p[2] = b;
    FN(clip_left)(p, &n, o, c);
    FN(clip_left)(p, &n, c, d);
    FN(clip_left)(p, &n, d, o);
    res = (REAL)fabs((double)FN(signed_area)(p, n));
    if (s1 * s2 == -1) res = -res;
    return res;
}

static void FN(reverse4)(PT* ps) {
    PT t = ps[0];
    ps[0] = ps[3];
    ps[3] = t;
    t = ps[1];
    ps[1] = ps[2];
    ps[2] = t;
}

/* IoU of two quadrilaterals given as 8 coordinates each (polyiou.cpp:91-133) */
REAL FN(oracle_iou_poly)(const REAL* pa, const REAL* qa) {
    PT p[6], q[6];
    int i, j;
    REAL inter = 0, a1, a2, uni;
    for (i = 0; i < 4; i++) {
        p[i].x = pa[2 * i];
        p[i].y = pa[2 * i + 1];
        q[i].x = qa[2 * i];
        q[i].y = qa[2 * i + 1];
    }
    if (FN(signed_area)(p, 4) < 0) FN(reverse4)(p);
    if (FN(signed_area)(q, 4) < 0) FN(reverse4)(q);
    p[4] = p[0];
    q[4] = q[0];
    for (i = 0; i < 4; i++)
        for (j = 0; j < 4; j++) inter += FN(tri_overlap)(p[i], p[i + 1], q[j], q[j + 1]);
    a1 = (REAL)fabs((double)FN(signed_area)(p, 4));
    a2 = (REAL)fabs((double)FN(signed_area)(q, 4));
    uni = a1 + a2 - inter;
    if (uni == 0) return (inter + 1) / (uni + 1);
    return inter / uni;
}

void FN(oracle_iou_poly_batch)(const REAL* p, const REAL* q, REAL* out, int n) {
    int i;
    for (i = 0; i < n; i++) out[i] = FN(oracle_iou_poly)(p + 8 * i, q + 8 * i);
}

/* Greedy NMS over boxes visited in `order` (descending score): box i is kept unless an already kept, earlier box
 * j has IoU(j, i) > thresh -- IoU(row = higher score, col = lower score), strict '>' as in the external kernel.
 * boxes are float32 [n][8] in both instantiations; the double one widens them (accuracy probe, not the reference).
 * Returns the number kept; keep[] receives indices into the INPUT order (= order[kept positions]). */
int FN(oracle_poly_nms)(const float* boxes, const int32_t* order, int n, float thresh, int32_t* keep) {
    int nk = 0, a, b, c;
    for (a = 0; a < n; a++) {
        int i = order[a];
        REAL bi[8];
        int dead = 0;
        for (c = 0; c < 8; c++) bi[c] = (REAL)boxes[8 * i + c];
        for (b = 0; b < nk && !dead; b++) {
            int j = keep[b];
            REAL bj[8];
            for (c = 0; c < 8; c++) bj[c] = (REAL)boxes[8 * j + c];
            if (FN(oracle_iou_poly)(bj, bi) > (REAL)thresh) dead = 1;
        }
        if (!dead) keep[nk++] = i;
    }
    return nk;
}
