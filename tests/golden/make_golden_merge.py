"""Golden vectors for the patch-merge polygon NMS, produced by the REFERENCE's own code. Run HERE (the
container that has /root/reference), never on the GPU box:

    python tests/golden/make_golden_merge.py

The function `py_cpu_nms_poly_fast` is taken from /root/reference/dafne/utils/ResultMerge_multi_process.py:61-122 by
reading that file at generation time and exec-ing the function's source (the module itself cannot be imported: it pulls
in the un-installed `polyiou` SWIG module and dota_utils). `polyiou.iou_poly` / `polyiou.VectorDouble` are bound to the
reference's tools/prepare_dota/polyiou.cpp compiled into oracle/_ref/libpolyiou_ref.so (oracle/Makefile). Nothing of
the reference is copied into this repository; only inputs and outputs are stored (tests/golden/patch_merge_nms.npz).
"""
import ctypes as C
import math
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF_FILE = "/root/reference/dafne/utils/ResultMerge_multi_process.py"

from oracle import postprocess as opost  # noqa: E402


def reference_function():
    src = open(REF_FILE).read()
    a = src.index("def py_cpu_nms_poly_fast(")
    b = src.index("\ndef ", a + 1)
    ref = opost.ref_lib()
    assert ref is not None, "run `make -C oracle` first (needs /root/reference)"
    ref.ref_iou_poly.restype = C.c_double
    dp = C.POINTER(C.c_double)

    def iou_poly(p, q):
        pa = (C.c_double * 8)(*p)
        qa = (C.c_double * 8)(*q)
        return ref.ref_iou_poly(C.cast(pa, dp), C.cast(qa, dp))

    polyiou = types.SimpleNamespace(VectorDouble=list, iou_poly=iou_poly)
    ns = {"np": np, "math": math, "polyiou": polyiou, "pdb": types.SimpleNamespace(set_trace=lambda: None)}
    exec(compile(src[a:b], REF_FILE, "exec"), ns)
    return ns["py_cpu_nms_poly_fast"]


def rot_rects64(rng, n, extent, wmin, wmax, aspect):
    cx, cy = rng.uniform(0, extent, n), rng.uniform(0, extent, n)
    w = rng.uniform(wmin, wmax, n)
    h = w / aspect
    a = rng.uniform(0, np.pi, n)
    dx = np.stack([-w, w, w, -w], 1) / 2
    dy = np.stack([-h, -h, h, h], 1) / 2
    x = cx[:, None] + dx * np.cos(a)[:, None] - dy * np.sin(a)[:, None]
    y = cy[:, None] + dx * np.sin(a)[:, None] + dy * np.cos(a)[:, None]
    return np.stack([x, y], 2).reshape(n, 8)


def cases():
    rng = np.random.default_rng(7)
    out = {}
    # (name, n, extent of the full image, box width range, aspect): from sparse to heavily overlapping
    for name, n, extent, wmin, wmax, asp in (("one", 1, 100, 20, 40, 3.0), ("sparse", 60, 4000, 20, 80, 3.0),
                                             ("ships", 700, 1500, 30, 160, 6.0), ("dense", 1500, 900, 20, 90, 3.0),
                                             ("dup", 300, 400, 20, 60, 2.0)):
        boxes = rot_rects64(rng, n, extent, wmin, wmax, asp)
        if name == "dup":  # the same object reported by overlapping patches: near-duplicates after the offset shift
            boxes[150:] = boxes[:150] + rng.normal(0, 1.5, (150, 8))
        # detections as mergesingle builds them: float(text) coordinates / rate, confidence with 3+ decimals
        boxes = np.round(boxes, 1)
        scores = rng.permutation(n) / n * 0.95 + 0.05 + rng.uniform(0, 1e-4, n)  # distinct
        out[name] = np.concatenate([boxes, scores[:, None]], 1).astype(np.float64)
    return out


def main():
    fn = reference_function()
    blob = {}
    for name, dets in cases().items():
        for thr in (0.1, 0.3):
            keep = np.asarray(fn(dets.copy(), thr), np.int64)
            blob[f"{name}_keep_{thr}"] = keep
            print(name, thr, len(dets), "->", len(keep))
        blob[f"{name}_dets"] = dets
    np.savez_compressed(os.path.join(HERE, "patch_merge_nms.npz"), **blob)


if __name__ == "__main__":
    main()
