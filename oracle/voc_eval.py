"""ORACLE (test infrastructure only -- the product path never imports this package).

CPU restatement of the reference's VOC-style average precision with polygon IoU:

  voc_ap(rec, prec, use_07_metric)                                  dafne/evaluation/voc_eval.py:7-38
  voc_eval(detpath, annopath, imagesetfile, classname, ...)         dafne/evaluation/voc_eval.py:41-224
  parse_gt(filename)                                                dafne/evaluation/dota_evaluation.py:73-109

The polygon IoU is the double instantiation of oracle/polyiou_oracle.c (pinned against the reference's polyiou.cpp). The
whole function is pinned against golden vectors produced by the reference's own voc_eval (tests/golden/
make_golden_voc.py -> tests/golden/voc_eval.npz).
"""
from __future__ import annotations

import numpy as np

from . import postprocess as _op


def voc_ap(rec, prec, use_07_metric=False):
    if use_07_metric:
        ap = 0.0
        for t in np.arange(0.0, 1.1, 0.1):
            p = 0 if np.sum(rec >= t) == 0 else np.max(prec[rec >= t])
            ap = ap + p / 11.0
        return ap
    mrec = np.concatenate(([0.0], rec, [1.0]))
    mpre = np.concatenate(([0.0], prec, [0.0]))
    for i in range(mpre.size - 1, 0, -1):
        mpre[i - 1] = np.maximum(mpre[i - 1], mpre[i])
    i = np.where(mrec[1:] != mrec[:-1])[0]
    return np.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1])


def parse_gt(filename):
    objects = []
    with open(filename, "r") as f:
        for line in f:
            s = line.strip().split(" ")
            if len(s) < 9:
                continue
            obj = {"name": s[8], "bbox": [float(v) for v in s[:8]]}
            if len(s) == 9:
                obj["difficult"] = 0
            elif len(s) == 10:
                obj["difficult"] = int(s[9])
            objects.append(obj)
    return objects


def match_detection(bb: np.ndarray, BBGT: np.ndarray):
    """voc_eval.py:133-186: (ovmax, jmax) of one detection against the ground truths of its image."""
    ovmax, jmax = -np.inf, -1
    if BBGT.size > 0:
        gx0, gy0 = BBGT[:, 0::2].min(1), BBGT[:, 1::2].min(1)
        gx1, gy1 = BBGT[:, 0::2].max(1), BBGT[:, 1::2].max(1)
        bx0, by0, bx1, by1 = bb[0::2].min(), bb[1::2].min(), bb[0::2].max(), bb[1::2].max()
        iw = np.maximum(np.minimum(gx1, bx1) - np.maximum(gx0, bx0) + 1.0, 0.0)
        ih = np.maximum(np.minimum(gy1, by1) - np.maximum(gy0, by0) + 1.0, 0.0)
        inters = iw * ih
        uni = (bx1 - bx0 + 1.0) * (by1 - by0 + 1.0) + (gx1 - gx0 + 1.0) * (gy1 - gy0 + 1.0) - inters
        keep = np.where(inters / uni > 0)[0]
        if keep.size:
            ov = _op.iou_poly_batch(np.ascontiguousarray(BBGT[keep]), np.ascontiguousarray(np.repeat(bb[None], keep.size, 0)),
                                    double=True)  # iou_poly(GT, bb)
            ovmax = ov.max()
            jmax = int(keep[int(np.argmax(ov))])
    return ovmax, jmax


def voc_eval(detpath, annopath, imagesetfile, classname, ovthresh=0.5, use_07_metric=False, parse_gt=parse_gt):
    with open(imagesetfile, "r") as f:
        imagenames = [x.strip() for x in f.readlines()]
    class_recs, npos = {}, 0
    for name in imagenames:
        R = [o for o in parse_gt(annopath.format(name)) if o["name"] == classname]
        bbox = np.array([x["bbox"] for x in R])
        difficult = np.array([x["difficult"] for x in R]).astype(bool)
        npos += int(np.sum(~difficult))
        class_recs[name] = {"bbox": bbox, "difficult": difficult, "det": [False] * len(R)}
    with open(detpath.format(classname), "r") as f:
        splitlines = [x.strip().split(" ") for x in f.readlines()]
    image_ids = [x[0] for x in splitlines]
    confidence = np.array([float(x[1]) for x in splitlines])
    BB = np.array([[float(z) for z in x[2:]] for x in splitlines])
    sorted_ind = np.argsort(-confidence)
    if BB.shape[0] > 0:
        BB = BB[sorted_ind, :]
    image_ids = [image_ids[x] for x in sorted_ind]
    nd = len(image_ids)
    tp, fp = np.zeros(nd), np.zeros(nd)
    for d in range(nd):
        R = class_recs[image_ids[d]]
        ovmax, jmax = match_detection(BB[d].astype(float), R["bbox"].astype(float))
        if ovmax > ovthresh:
            if not R["difficult"][jmax]:
                if not R["det"][jmax]:
                    tp[d] = 1.0
                    R["det"][jmax] = 1
                else:
                    fp[d] = 1.0
        else:
            fp[d] = 1.0
    fp, tp = np.cumsum(fp), np.cumsum(tp)
    rec = tp / float(npos)
    prec = tp / np.maximum(tp + fp, np.finfo(np.float64).eps)
    return rec, prec, voc_ap(rec, prec, use_07_metric)
