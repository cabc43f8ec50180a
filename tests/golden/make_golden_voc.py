"""Golden vectors for the polygon-IoU VOC evaluation, produced by the REFERENCE's own code. Run HERE (the
container that has /root/reference), never on the GPU box:

    python tests/golden/make_golden_voc.py

`voc_ap` and `voc_eval` are taken from /root/reference/dafne/evaluation/voc_eval.py:7-224 and `parse_gt` from
/root/reference/dafne/evaluation/dota_evaluation.py:73-109 by reading those files at generation time and exec-ing the
functions' source (the modules themselves cannot be imported: detectron2, the `polyiou` SWIG module). `polyiou` is bound
to the reference's tools/prepare_dota/polyiou.cpp compiled into oracle/_ref/libpolyiou_ref.so; `np.bool`, which the
reference still uses (voc_eval.py:96) and NumPy 2 removed, is mapped to `bool`. Nothing of the reference is copied into
this repository; only the synthetic inputs and the reference's outputs are stored (tests/golden/voc_eval.npz).
"""
import ctypes as C
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF_VOC = "/root/reference/dafne/evaluation/voc_eval.py"
REF_DOTA = "/root/reference/dafne/evaluation/dota_evaluation.py"

from oracle import postprocess as opost  # noqa: E402
from tests.golden.make_golden_merge import rot_rects64  # noqa: E402


class _Np:
    """numpy with the removed alias the reference still uses."""
    bool = bool

    def __getattr__(self, name):
        return getattr(np, name)


def reference_functions():
    ref = opost.ref_lib()
    assert ref is not None, "run `make -C oracle` first (needs /root/reference)"
    ref.ref_iou_poly.restype = C.c_double
    dp = C.POINTER(C.c_double)

    def iou_poly(p, q):
        return ref.ref_iou_poly(C.cast((C.c_double * 8)(*p), dp), C.cast((C.c_double * 8)(*q), dp))

    polyiou = types.SimpleNamespace(VectorDouble=list, iou_poly=iou_poly)
    ns = {"np": _Np(), "polyiou": polyiou}
    src = open(REF_VOC).read()
    a = src.index("def voc_ap(")
    exec(compile(src[a:], REF_VOC, "exec"), ns)
    dsrc = open(REF_DOTA).read()
    a = dsrc.index("def parse_gt(")
    b = dsrc.index("\ndef ", a + 1)
    exec(compile(dsrc[a:b], REF_DOTA, "exec"), ns)
    return ns["voc_eval"], ns["parse_gt"]


def synth(seed, n_img, gt_per_img, det_per_img, classes=("plane", "ship")):
    """Ground truth + detections in the reference's file formats, as lists of lines."""
    rng = np.random.default_rng(seed)
    gt_lines = {}
    det_lines = {c: [] for c in classes}
    for i in range(n_img):
        name = f"P{i:04d}"
        lines = []
        gts = {c: [] for c in classes}
        for _ in range(int(rng.integers(0, gt_per_img + 1))):
            c = classes[int(rng.integers(0, len(classes)))]
            box = np.round(rot_rects64(rng, 1, 800, 25, 120, float(rng.uniform(1.5, 6.0)))[0], 1)
            diff = int(rng.random() < 0.15)
            lines.append(" ".join(str(v) for v in box) + f" {c} {diff}")
            gts[c].append(box)
        gt_lines[name] = lines
        for c in classes:
            for g in gts[c]:  # detections near most ground truths (some twice), plus clutter
                for _ in range(int(rng.integers(0, 3))):
                    b = np.round(g + rng.normal(0, 3.0, 8), 1)
                    det_lines[c].append(f"{name} {rng.uniform(0.05, 1.0):.6f} " + " ".join(str(v) for v in b))
            for _ in range(int(rng.integers(0, det_per_img + 1))):
                b = np.round(rot_rects64(rng, 1, 800, 25, 120, 3.0)[0], 1)
                det_lines[c].append(f"{name} {rng.uniform(0.05, 0.6):.6f} " + " ".join(str(v) for v in b))
    for c in classes:
        rng.shuffle(det_lines[c])
    return gt_lines, det_lines


def write_case(tmp, gt_lines, det_lines):
    os.makedirs(os.path.join(tmp, "gt"), exist_ok=True)
    for name, lines in gt_lines.items():
        with open(os.path.join(tmp, "gt", name + ".txt"), "w") as f:
            f.write("\n".join(lines) + ("\n" if lines else ""))
    with open(os.path.join(tmp, "imageset.txt"), "w") as f:
        f.write("\n".join(gt_lines) + "\n")
    for c, lines in det_lines.items():
        with open(os.path.join(tmp, f"Task1_{c}.txt"), "w") as f:
            f.write("\n".join(lines) + "\n")
    return os.path.join(tmp, "Task1_{:s}.txt"), os.path.join(tmp, "gt", "{:s}.txt"), os.path.join(tmp, "imageset.txt")


def main():
    voc_eval, parse_gt = reference_functions()
    blob = {}
    for tag, seed, n_img, gpi, dpi in (("small", 1, 6, 5, 4), ("medium", 2, 40, 12, 10)):
        gt_lines, det_lines = synth(seed, n_img, gpi, dpi)
        blob[f"{tag}_gt_names"] = np.array(list(gt_lines))
        blob[f"{tag}_gt_lines"] = np.array(["\n".join(v) for v in gt_lines.values()])
        with tempfile.TemporaryDirectory() as tmp:
            detpath, annopath, imageset = write_case(tmp, gt_lines, det_lines)
            for c in det_lines:
                blob[f"{tag}_det_{c}"] = np.array(det_lines[c])
                for m07 in (True, False):
                    rec, prec, ap, _ = voc_eval(detpath, annopath, imageset, c, ovthresh=0.5, use_07_metric=m07,
                                                parse_gt=parse_gt)
                    blob[f"{tag}_{c}_rec"], blob[f"{tag}_{c}_prec"] = rec, prec
                    blob[f"{tag}_{c}_ap{'07' if m07 else '12'}"] = np.float64(ap)
                    print(tag, c, "07" if m07 else "12", len(det_lines[c]), "dets -> ap", ap)
    np.savez_compressed(os.path.join(HERE, "voc_eval.npz"), **blob)


if __name__ == "__main__":
    main()
