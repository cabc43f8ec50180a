"""ORACLE (test infrastructure only -- the product path never imports this package, and this file imports nothing
from dafne_b200).

CPU restatement of the merge half of the reference's test-time augmentation (dafne/modeling/tta.py:232-268): the
per-copy corners go through `tfm.inverse().apply_coords` in numpy float32 exactly like the reference does on the host,
the union through ml_nms (class offsets, polygon NMS) and the post-NMS top-k of select_over_all_levels
(dafne/modeling/dafne/dafne_outputs.py:907-925).

The transforms are restated HERE, independently of the product's classes, from the published behaviour of what the
reference's mapper instantiates (dafne/modeling/tta.py:69-135; detectron2 v0.5 / fvcore, not installable offline --
"parity unpinned" for this (f)-row):
  ResizeShortestEdge.get_transform   scale = size / min(h, w); cap the long edge at max_size; int(x + 0.5)
  ResizeTransform.apply_coords       x * (new_w * 1.0 / w), y * (new_h * 1.0 / h)      inverse: swap old / new
  HFlipTransform.apply_coords        x = width - x                                     inverse: itself
  VFlipTransform.apply_coords        y = height - y                                    inverse: itself
  TransformList.inverse              inverses in reverse order
A transform chain is described by plain tuples: ("resize", h, w, new_h, new_w) | ("hflip", width) | ("vflip", height).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

from . import postprocess as _op

F32 = np.float32


def shortest_edge_size(h: int, w: int, size: int, max_size: int) -> Tuple[int, int]:
    """detectron2 v0.5 ResizeShortestEdge: the (new_h, new_w) of an h x w image."""
    scale = size * 1.0 / min(h, w)
    if h < w:
        newh, neww = size, scale * w
    else:
        newh, neww = scale * h, size
    if max(newh, neww) > max_size:
        scale = max_size * 1.0 / max(newh, neww)
        newh, neww = newh * scale, neww * scale
    return int(newh + 0.5), int(neww + 0.5)


def chain_of_copy(h: int, w: int, min_size: int, max_size: int, flip: str = "", orig_hw=None) -> List[tuple]:
    """The transform chain the mapper builds for one augmented copy of an h x w input image (tta.py:69-135, detectron2
    DatasetMapperTTA): `pre_tfm` from the ORIGINAL image (dataset_dict["height"], ["width"]) to the input tensor when
    the two differ, the resize, then an optional flip of the RESIZED image."""
    chain = []
    if orig_hw is not None and tuple(orig_hw) != (h, w):
        chain.append(("resize", orig_hw[0], orig_hw[1], h, w))
    nh, nw = shortest_edge_size(h, w, min_size, max_size)
    chain.append(("resize", h, w, nh, nw))
    if flip == "h":
        chain.append(("hflip", nw))
    elif flip == "v":
        chain.append(("vflip", nh))
    return chain


def inverse_chain(chain: Sequence[tuple]) -> List[tuple]:
    out = []
    for t in reversed(list(chain)):
        if t[0] == "resize":
            _, h, w, nh, nw = t
            out.append(("resize", nh, nw, h, w))
        elif t[0] in ("hflip", "vflip", "noop"):
            out.append(t)
        else:
            raise ValueError(f"unknown transform {t!r}")
    return out


def apply_coords(chain: Sequence[tuple], coords: np.ndarray) -> np.ndarray:
    """[n, 2] float32 (x, y) through the chain, in numpy float32 with python-float scalars like fvcore / detectron2."""
    coords = np.array(coords, dtype=np.float32, copy=True)
    for t in chain:
        if t[0] == "resize":
            _, h, w, nh, nw = t
            coords[:, 0] = coords[:, 0] * (nw * 1.0 / w)
            coords[:, 1] = coords[:, 1] * (nh * 1.0 / h)
        elif t[0] == "hflip":
            coords[:, 0] = t[1] - coords[:, 0]
        elif t[0] == "vflip":
            coords[:, 1] = t[1] - coords[:, 1]
        elif t[0] != "noop":
            raise ValueError(f"unknown transform {t!r}")
    return coords


def merge_detections(corners: Sequence[np.ndarray], scores: Sequence[np.ndarray], classes: Sequence[np.ndarray],
                     tfms: Sequence, nms_thresh: float = 0.1, post_nms_topk: int = 1000,
                     vehicle_merge: bool = True):
    """corners[k]: [n_k, 8] float32 in the coordinates of augmented copy k; tfms[k]: the copy's transform chain as
    tuples (see the module header). Returns (corners, scores, classes, index into the concatenated input) of the merged
    result, best score first."""
    orig: List[np.ndarray] = []
    for c, t in zip(corners, tfms):
        n = c.shape[0]
        pts = np.ascontiguousarray(c, dtype=np.float32).reshape(-1, 2)
        orig.append(apply_coords(inverse_chain(t), pts).reshape(n, 8).astype(np.float32))
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
