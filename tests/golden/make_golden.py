"""Generates the committed golden fixtures. Run HERE (the container that has /root/reference), never on the GPU box:

    python tests/golden/make_golden.py

  sort_corners.npz   inputs + outputs of the REFERENCE's own dafne/utils/sort_corners.py::sort_quadrilateral,
                     imported from /root/reference by path (it only needs torch)
  polyiou_ref.npz    quad pairs + IoU from the REFERENCE's tools/prepare_dota/polyiou.cpp (compiled into
                     oracle/_ref/libpolyiou_ref.so by oracle/Makefile), double precision
  postprocess_*.npz  small head outputs + the ORACLE's post-processing result (oracle-generated regression fixture;
                     the reference has no test for this boundary and cannot run here: detectron2 / poly_nms missing)
"""
import ctypes as C
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import postprocess as opost  # noqa: E402


def rot_rects(rng, n, w=48, h=16, center=(0, 1024), jitter=0.0):
    cx = rng.uniform(*center, n)
    cy = rng.uniform(*center, n)
    ww = w * rng.uniform(0.5, 2.0, n)
    hh = h * rng.uniform(0.5, 2.0, n)
    a = rng.uniform(0, np.pi, n)
    dx = np.stack([-ww, ww, ww, -ww], 1) / 2
    dy = np.stack([-hh, -hh, hh, hh], 1) / 2
    x = cx[:, None] + dx * np.cos(a)[:, None] - dy * np.sin(a)[:, None]
    y = cy[:, None] + dx * np.sin(a)[:, None] + dy * np.cos(a)[:, None]
    q = np.stack([x, y], 2).reshape(n, 8)
    return (q + rng.normal(0, jitter, q.shape)).astype(np.float32)


def sort_corner_inputs():
    rng = np.random.default_rng(0)
    parts = [rng.normal(0, 50, (600, 8)), rot_rects(rng, 600), rot_rects(rng, 300, jitter=6.0)]
    # shuffled vertex orders of rectangles
    r = rot_rects(rng, 300).reshape(-1, 4, 2)
    perm = np.stack([rng.permutation(4) for _ in range(300)])
    parts.append(r[np.arange(300)[:, None], perm].reshape(-1, 8))
    # ties in x (leftmost ambiguity), integer grids
    parts.append(rng.integers(0, 4, (400, 8)).astype(np.float32))
    # degenerate: repeated points, collinear points, all equal
    d = rng.normal(0, 10, (100, 8))
    d[:, 2:4] = d[:, 0:2]
    parts.append(d)
    t = rng.uniform(0, 1, (100, 4))
    parts.append(np.stack([t * 10, t * 3 + 1], 2).reshape(100, 8))
    parts.append(np.tile(rng.normal(0, 5, (20, 2)), (1, 4)))
    return np.concatenate(parts).astype(np.float32)


def main():
    # ---- sort_corners from the reference file itself
    spec = importlib.util.spec_from_file_location("ref_sort_corners", os.path.join(REF, "dafne/utils/sort_corners.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    quads = sort_corner_inputs()
    out = mod.sort_quadrilateral(torch.from_numpy(quads)).numpy()
    np.savez_compressed(os.path.join(HERE, "sort_corners.npz"), quads=quads, sorted=out)
    print("sort_corners:", quads.shape)

    # ---- polygon IoU from the reference's polyiou.cpp
    ref = opost.ref_lib()
    assert ref is not None, "run `make -C oracle` first (needs /root/reference)"
    rng = np.random.default_rng(1)
    n = 2000
    p = rot_rects(rng, n, center=(0, 200))
    q = rot_rects(rng, n, center=(0, 200))
    q[:400] = p[:400] + rng.normal(0, 4, (400, 8)).astype(np.float32)  # heavy overlaps
    q[400:450] = p[400:450]  # identical
    q[450:500] = p[450:500].reshape(-1, 4, 2)[:, ::-1].reshape(-1, 8)  # same quad, opposite orientation
    p[500:520, 2:] = np.tile(p[500:520, :2], (1, 3))  # zero-area (point) polygons
    q[500:520] = p[500:520]
    known_p = np.array([[0, 0, 1, 0, 1, 1, 0, 1], [0, 0, 1, 0, 1, 1, 0, 1],
                        [686, 2976, 709, 2976, 724, 2976, 701, 2976]], np.float64)
    known_q = np.array([[0.5, 0.5, 1.5, 0.5, 1.5, 1.5, 0.5, 1.5], [0.5, -0.5, 1.5, 0.5, 0.5, 1.5, -0.5, 0.5],
                        [686, 2976, 709, 2976, 724, 2976, 701, 2976]], np.float64)
    pd = np.concatenate([p.astype(np.float64), known_p])
    qd = np.concatenate([q.astype(np.float64), known_q])
    out = np.empty(len(pd))
    dp = C.POINTER(C.c_double)
    ref.ref_iou_poly_batch(pd.ctypes.data_as(dp), qd.ctypes.data_as(dp), out.ctypes.data_as(dp), len(pd))
    np.savez_compressed(os.path.join(HERE, "polyiou_ref.npz"), p=pd, q=qd, iou=out)
    print("polyiou_ref:", pd.shape, "known answers:", out[-3:])

    # ---- oracle post-processing regression fixtures
    for tag, C_, sort_c, twc, seed in (("c15_sort", 15, True, False, 3), ("c15_ctr", 15, True, True, 4),
                                       ("c1_nosort", 1, False, False, 5)):
        rng = np.random.default_rng(seed)
        N, Hs, Ws = 2, [24, 12, 6, 3, 2], [32, 16, 8, 4, 2]
        strides = [8, 16, 32, 64, 128]
        logits, reg, ctr = [], [], []
        for h, w in zip(Hs, Ws):
            bias = -6.2 if twc else -3.4
            logits.append((rng.normal(bias, 1.2, (N, C_, h, w))).astype(np.float32))
            base = np.array([-3, -1, 3, -1, 3, 1, -3, 1], np.float32).reshape(1, 8, 1, 1)
            reg.append((base + rng.normal(0, 1.0, (N, 8, h, w))).astype(np.float32))
            ctr.append(rng.normal(0, 1.0, (N, 1, h, w)).astype(np.float32))
        sizes = [(Hs[0] * 8, Ws[0] * 8), (Hs[0] * 8 - 20, Ws[0] * 8 - 12)]
        osz = [(Hs[0] * 16, Ws[0] * 16), sizes[1]]
        res = opost.postprocess(logits, reg, ctr, strides, sizes, osz, pre_nms_topk=60, post_nms_topk=40,
                                sort_corners=sort_c, thresh_with_ctr=twc)
        blob = {}
        for l in range(5):
            blob[f"logits{l}"], blob[f"reg{l}"], blob[f"ctr{l}"] = logits[l], reg[l], ctr[l]
        for i, r in enumerate(res):
            for k, v in r.items():
                blob[f"out{i}_{k}"] = np.asarray(v)
        blob["sizes"] = np.array(sizes)
        blob["osz"] = np.array(osz)
        blob["meta"] = np.array([C_, int(sort_c), int(twc), 60, 40])
        np.savez_compressed(os.path.join(HERE, f"postprocess_{tag}.npz"), **blob)
        print("postprocess", tag, [len(r["scores"]) for r in res])


if __name__ == "__main__":
    main()
