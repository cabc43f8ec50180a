"""ORACLE (test infrastructure only -- the product path never imports this package).

CPU restatement of the merge half of the reference's test-time augmentation (dafne/modeling/tta.py:232-268): the
per-copy corners go through `tfm.inverse().apply_coords` in numpy float32 exactly like the reference does on the host,
the union through ml_nms (class offsets, polygon NMS) and the post-NMS top-k of select_over_all_levels
(dafne/modeling/dafne/dafne_outputs.py:907-925). "Parity unpinned": detectron2's transform classes cannot be installed
here; the arithmetic of apply_coords (scale by new/old, width - x, height - y) is restated in dafne_b200/tta.py and
exercised here through its numpy path.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

from . import postprocess as _op

F32 = np.float32


def merge_detections(corners: Sequence[np.ndarray], scores: Sequence[np.ndarray], classes: Sequence[np.ndarray],
                     tfms: Sequence, nms_thresh: float = 0.1, post_nms_topk: int = 1000,
                     vehicle_merge: bool = True):
    """corners[k]: [n_k, 8] float32 in the coordinates of augmented copy k; tfms[k]: its transform (with inverse()).
    Returns (corners, scores, classes, index into the concatenated input) of the merged result, best score first."""
    orig: List[np.ndarray] = []
    for c, t in zip(corners, tfms):
        n = c.shape[0]
        pts = np.ascontiguousarray(c, dtype=np.float32).reshape(-1, 2).copy()
        orig.append(t.inverse().apply_coords(pts).reshape(n, 8).astype(np.float32))
    poly = np.concatenate(orig, 0)
    sc = np.concatenate([np.asarray(s, np.float32) for s in scores], 0)
    cl = np.concatenate([np.asarray(c, np.int64) for c in classes], 0)
    n = len(sc)
    if n == 0:
        return poly, sc, cl, np.zeros(0, np.int64)
    order = np.lexsort((np.arange(n), -sc.astype(np.float64)))  # descending score, ties by ascending index
    idx = cl.copy()
    if vehicle_merge:
        idx[idx == 5] = 4
    span = (poly.max() - poly.min()) + F32(1.0)
    boxes = poly + (idx.astype(np.float32) * span)[:, None]
    keep = order[_op.greedy_nms(boxes[order], nms_thresh)]
    if post_nms_topk > 0 and len(keep) > post_nms_topk:
        kth = np.sort(sc[keep])[len(keep) - post_nms_topk]
        keep = keep[sc[keep] >= kth]
    return poly[keep], sc[keep], cl[keep], keep
