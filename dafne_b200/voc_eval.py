"""VOC-style average precision with polygon IoU on the GPU: the reference's dafne/evaluation/voc_eval.py
with the same function names, arguments and return values.

  voc_ap(rec, prec, use_07_metric=False)                                               voc_eval.py:7-38
  voc_eval(detpath, annopath, imagesetfile, classname, ovthresh=0.5, use_07_metric=False, parse_gt=None)
                                                                                        voc_eval.py:41-224
  parse_gt(filename)                                                                    dota_evaluation.py:73-109

What moves to the device is the O(detections x ground truths) part: for every detection the best polygon IoU against the
ground truths of its image (`dafne_voc_match_f64_host`, double precision, the SWIG `polyiou.iou_poly` arithmetic). File
parsing, the sequential true/false-positive assignment and the AP integral stay on the host -- they are O(detections).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from .merge import _device


def voc_ap(rec, prec, use_07_metric=False):
    if use_07_metric:  # 11 point metric
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


def match_detections(BB: np.ndarray, det_image: np.ndarray, gts: np.ndarray, gt_offsets: np.ndarray):
    """(ovmax [nd] float64, jmax [nd] int32) for detections BB [nd, 8] whose images are det_image [nd]."""
    nd = BB.shape[0]
    ovmax = np.full(nd, -np.inf, np.float64)
    jmax = np.full(nd, -1, np.int32)
    if nd == 0:
        return ovmax, jmax
    BB = np.ascontiguousarray(BB, np.float64)
    det_image = np.ascontiguousarray(det_image, np.int32)
    gts = np.ascontiguousarray(gts, np.float64).reshape(-1, 8)
    gt_offsets = np.ascontiguousarray(gt_offsets, np.int32)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    _capi.check(_capi.lib().dafne_voc_match_f64_host(
        BB.ctypes.data_as(dp), det_image.ctypes.data_as(ip), nd, gts.ctypes.data_as(dp), gt_offsets.ctypes.data_as(ip),
        len(gt_offsets) - 1, _device(), ovmax.ctypes.data_as(dp), jmax.ctypes.data_as(ip)), "dafne_voc_match_f64_host")
    return ovmax, jmax


def voc_eval(detpath, annopath, imagesetfile, classname, ovthresh=0.5, use_07_metric=False, parse_gt=parse_gt):
    with open(imagesetfile, "r") as f:
        imagenames = [x.strip() for x in f.readlines()]
    class_recs, npos = {}, 0
    gt_rows, gt_offsets, image_index = [], [0], {}
    for k, name in enumerate(imagenames):
        R = [o for o in parse_gt(annopath.format(name)) if o["name"] == classname]
        bbox = np.array([x["bbox"] for x in R], np.float64).reshape(-1, 8)
        difficult = np.array([x["difficult"] for x in R]).astype(bool)
        npos += int(np.sum(~difficult))
        class_recs[name] = {"bbox": bbox, "difficult": difficult, "det": [False] * len(R)}
        gt_rows.append(bbox)
        gt_offsets.append(gt_offsets[-1] + len(R))
        image_index[name] = k
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
    # device: the polygon-IoU matching of all detections at once (voc_eval.py:133-186)
    det_image = np.array([image_index[i] for i in image_ids], np.int32)
    ovmaxs, jmaxs = match_detections(BB.reshape(-1, 8), det_image,
                                     np.concatenate(gt_rows, 0) if gt_rows else np.zeros((0, 8)), np.array(gt_offsets))
    tp, fp = np.zeros(nd), np.zeros(nd)
    data_scores_overlap = []
    for d in range(nd):  # voc_eval.py:187-206
        R = class_recs[image_ids[d]]
        ovmax, jmax = ovmaxs[d], int(jmaxs[d])
        if ovmax > ovthresh:
            if not R["difficult"][jmax]:
                if not R["det"][jmax]:
                    tp[d] = 1.0
                    R["det"][jmax] = 1
                    data_scores_overlap.append([confidence[d], ovmax, 1, classname])
                else:
                    fp[d] = 1.0
                    data_scores_overlap.append([confidence[d], ovmax, 0, classname])
        else:
            fp[d] = 1.0
    fp, tp = np.cumsum(fp), np.cumsum(tp)
    rec = tp / float(npos)
    prec = tp / np.maximum(tp + fp, np.finfo(np.float64).eps)
    return rec, prec, voc_ap(rec, prec, use_07_metric), data_scores_overlap
