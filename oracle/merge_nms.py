"""ORACLE (test infrastructure only -- the product path never imports this package).

CPU restatement of the polygon NMS of the reference's patch-merge step:

  py_cpu_nms_poly_fast(dets, thresh)   dafne/utils/ResultMerge_multi_process.py:61-122
  nmsbynamedict / poly2origpoly        dafne/utils/ResultMerge_multi_process.py:155-181

The IoU is the double-precision instantiation of oracle/polyiou_oracle.c, which is pinned against the reference's own
tools/prepare_dota/polyiou.cpp (tests/test_oracle_cpu.py). The whole function is pinned against golden vectors produced
by the reference's own py_cpu_nms_poly_fast (tests/golden/make_golden_merge.py -> tests/golden/patch_merge_nms.npz).

Convention the reference leaves open (numpy `argsort()[::-1]` does not define the order of equal scores): descending
score, ties by ascending input index.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np

from . import postprocess as _op


def py_cpu_nms_poly_fast(dets: np.ndarray, thresh: float) -> List[int]:
    dets = np.asarray(dets, np.float64)
    n = dets.shape[0]
    if n == 0:
        return []
    obbs = np.ascontiguousarray(dets[:, 0:8])
    x1, y1 = obbs[:, 0::2].min(1), obbs[:, 1::2].min(1)  # :63-66
    x2, y2 = obbs[:, 0::2].max(1), obbs[:, 1::2].max(1)
    scores = dets[:, 8]
    areas = (x2 - x1 + 1) * (y2 - y1 + 1)  # :68
    order = np.lexsort((np.arange(n), -scores))
    keep: List[int] = []
    while order.size > 0:
        i = order[0]
        keep.append(int(i))
        rest = order[1:]
        w = np.maximum(0.0, np.minimum(x2[i], x2[rest]) - np.maximum(x1[i], x1[rest]))  # :92-98 (no "+ 1" here)
        h = np.maximum(0.0, np.minimum(y2[i], y2[rest]) - np.maximum(y1[i], y1[rest]))
        hbb_inter = w * h
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = hbb_inter / (areas[i] + areas[rest] - hbb_inter)  # :100
        cand = np.nonzero(ovr > 0)[0]  # :102 -- only hbox-overlapping pairs get the polygon IoU
        if cand.size:
            p = np.ascontiguousarray(np.repeat(obbs[i][None], cand.size, 0))
            ovr[cand] = _op.iou_poly_batch(p, np.ascontiguousarray(obbs[rest[cand]]), double=True)  # :105-106
        order = rest[ovr <= thresh]  # :117,122 (a NaN overlap drops the box, as `<=` does in the reference)
    return keep


def poly2origpoly(poly, x, y, rate):  # :174-181
    out = []
    for i in range(len(poly) // 2):
        out.append(float(poly[i * 2] + x) / float(rate))
        out.append(float(poly[i * 2 + 1] + y) / float(rate))
    return out


def nmsbynamedict(nameboxdict: Dict[str, list], thresh: float) -> Dict[str, list]:  # :155-173
    out = {}
    for name, boxes in nameboxdict.items():
        keep = py_cpu_nms_poly_fast(np.array(boxes, np.float64).reshape(-1, 9), thresh)
        out[name] = [boxes[k] for k in keep]
    return out
