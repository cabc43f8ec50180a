"""Patch-merge step of the DOTA evaluation on the GPU: the host-side mirror of the reference's
dafne/utils/ResultMerge_multi_process.py for the part that costs time there -- the per-(class, full image) polygon NMS
in double precision. Same function names and argument meaning as the reference:

  py_cpu_nms_poly_fast(dets, thresh)          ResultMerge_multi_process.py:61-122   -> dafne_poly_nms_f64_host
  nmsbynamedict(nameboxdict, nms, thresh)     :155-173  (all images of the dict in ONE device call)
  poly2origpoly(poly, x, y, rate)             :174-181
  mergesingle(dstpath, nms, fullname)         :183-221  (Task1 result file of one class -> merged result file)

The name `py_cpu_nms_poly_fast` is kept so that `mergebase(srcpath, dstpath, py_cpu_nms_poly_fast)` call sites read
like the reference's (ResultMerge_multi_process.py:252-254); nothing here runs the NMS arithmetic on the CPU.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from typing import Callable, Dict, List, Sequence

import numpy as np

from . import _capi

nms_thresh = 0.1  # ResultMerge_multi_process.py:22


def _device() -> int:
    import torch

    if not torch.cuda.is_available():
        raise _capi.DafneError("dafne_b200.merge needs a CUDA device: the CUDA kernels are the only execution path")
    return torch.cuda.current_device()


def py_cpu_nms_poly_fast(dets, thresh: float) -> List[int]:
    """dets: [n, 9] = 8 polygon coordinates + score (float64). Returns the kept indices, best score first."""
    dets = np.ascontiguousarray(np.asarray(dets, dtype=np.float64).reshape(-1, 9))
    n = dets.shape[0]
    if n == 0:
        return []
    keep = (C.c_int32 * n)()
    num = C.c_int32(0)
    _capi.check(_capi.lib().dafne_poly_nms_f64_host(dets.ctypes.data_as(C.POINTER(C.c_double)), n, float(thresh),
                                                    _device(), keep, C.byref(num)), "dafne_poly_nms_f64_host")
    return list(keep[: num.value])


def nms_many(det_lists: Sequence[np.ndarray], thresh: float) -> List[List[int]]:
    """py_cpu_nms_poly_fast of every list in one device call (one problem per list)."""
    arrs = [np.asarray(d, dtype=np.float64).reshape(-1, 9) for d in det_lists]
    if not arrs:
        return []
    offsets = np.zeros(len(arrs) + 1, np.int32)
    offsets[1:] = np.cumsum([a.shape[0] for a in arrs])
    total = int(offsets[-1])
    if total == 0:
        return [[] for _ in arrs]
    dets = np.ascontiguousarray(np.concatenate(arrs, 0))
    keep = np.zeros(total, np.int32)
    nkeep = np.zeros(len(arrs), np.int32)
    _capi.check(_capi.lib().dafne_poly_nms_f64_batch_host(
        dets.ctypes.data_as(C.POINTER(C.c_double)), offsets.ctypes.data_as(C.POINTER(C.c_int32)), len(arrs),
        float(thresh), _device(), keep.ctypes.data_as(C.POINTER(C.c_int32)),
        nkeep.ctypes.data_as(C.POINTER(C.c_int32))), "dafne_poly_nms_f64_batch_host")
    return [keep[offsets[p]: offsets[p] + nkeep[p]].tolist() for p in range(len(arrs))]


def nmsbynamedict(nameboxdict: Dict[str, list], nms: Callable, thresh: float) -> Dict[str, list]:
    """Reference signature; when `nms` is this module's py_cpu_nms_poly_fast all images go to the device together."""
    names = list(nameboxdict)
    if nms is py_cpu_nms_poly_fast:
        keeps = nms_many([np.array(nameboxdict[k], np.float64) for k in names], thresh)
    else:
        keeps = [nms(np.array(nameboxdict[k]), thresh) for k in names]
    return {k: [nameboxdict[k][i] for i in keep] for k, keep in zip(names, keeps)}


def poly2origpoly(poly, x, y, rate):
    origpoly = []
    for i in range(int(len(poly) / 2)):
        origpoly.append(float(poly[i * 2] + x) / float(rate))
        origpoly.append(float(poly[i * 2 + 1] + y) / float(rate))
    return origpoly


_XY = re.compile(r"__\d+___\d+")
_RATE = re.compile(r"__([\d+\.]+)__\d+___")


def merge_lines(lines: Sequence[str], nms: Callable = py_cpu_nms_poly_fast, thresh: float = nms_thresh) -> List[str]:
    """Task1 lines `<patch name> <confidence> <8 coordinates>` of one class -> merged lines per full image. Patch names
    carry the offset and the scale, e.g. P0006__1__0___824 (ResultMerge_multi_process.py:190-206)."""
    nameboxdict: Dict[str, list] = {}
    for line in lines:
        parts = line.strip().split(" ")
        if len(parts) < 10:
            continue
        subname = parts[0]
        oriname = subname.split("__")[0]
        x, y = (int(v) for v in re.findall(r"\d+", _XY.findall(subname)[0])[:2])
        rate = _RATE.findall(subname)[0]
        det = poly2origpoly(list(map(float, parts[2:10])), x, y, rate)
        det.append(float(parts[1]))
        nameboxdict.setdefault(oriname, []).append(det)
    merged = nmsbynamedict(nameboxdict, nms, thresh)
    out = []
    for imgname, dets in merged.items():
        for det in dets:
            out.append(imgname + " " + str(det[-1]) + " " + " ".join(map(str, det[0:-1])))
    return out


def mergesingle(dstpath: str, nms: Callable, fullname: str) -> None:
    name = os.path.splitext(os.path.basename(fullname))[0]
    with open(fullname, "r") as f_in:
        lines = f_in.readlines()
    with open(os.path.join(dstpath, name + ".txt"), "w") as f_out:
        for line in merge_lines(lines, nms, nms_thresh):
            f_out.write(line + "\n")
