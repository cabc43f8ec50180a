"""VOC AP with polygon IoU: oracle and device against golden vectors produced by the reference's own
voc_eval (tests/golden/make_golden_voc.py)."""
import os

import numpy as np
import pytest

from oracle import voc_eval as ovoc

GOLD = os.path.join(os.path.dirname(__file__), "golden", "voc_eval.npz")


def _write(tmp, g, tag):
    os.makedirs(os.path.join(tmp, "gt"))
    for n, l in zip(g[f"{tag}_gt_names"], g[f"{tag}_gt_lines"]):
        with open(os.path.join(tmp, "gt", str(n) + ".txt"), "w") as f:
            f.write(str(l) + ("\n" if str(l) else ""))
    with open(os.path.join(tmp, "imageset.txt"), "w") as f:
        f.write("\n".join(map(str, g[f"{tag}_gt_names"])) + "\n")
    for c in ("plane", "ship"):
        with open(os.path.join(tmp, f"Task1_{c}.txt"), "w") as f:
            f.write("\n".join(map(str, g[f"{tag}_det_{c}"])) + "\n")
    return os.path.join(tmp, "Task1_{:s}.txt"), os.path.join(tmp, "gt", "{:s}.txt"), os.path.join(tmp, "imageset.txt")


def _check(fn, tmp_path, tag):
    g = np.load(GOLD)
    d, a, i = _write(str(tmp_path), g, tag)
    for c in ("plane", "ship"):
        for m07 in (True, False):
            out = fn(d, a, i, c, 0.5, m07)
            assert np.array_equal(out[0], g[f"{tag}_{c}_rec"]) and np.array_equal(out[1], g[f"{tag}_{c}_prec"])
            assert out[2] == g[f"{tag}_{c}_ap{'07' if m07 else '12'}"]


@pytest.mark.parametrize("tag", ["small", "medium"])
def test_oracle_matches_reference_golden(tmp_path, tag):
    _check(ovoc.voc_eval, tmp_path, tag)


def test_voc_ap_known_values():
    rec = np.array([0.1, 0.2, 0.2, 0.5, 1.0])
    prec = np.array([1.0, 1.0, 0.66, 0.5, 0.4])
    assert abs(ovoc.voc_ap(rec, prec, False) - (0.1 + 0.1 + 0.3 * 0.5 + 0.5 * 0.4)) < 1e-12
    assert abs(ovoc.voc_ap(rec, prec, True) - (3 * 1.0 + 3 * 0.5 + 5 * 0.4) / 11.0) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["small", "medium"])
def test_device_matches_reference_golden(tmp_path, tag):
    from dafne_b200 import voc_eval as dvoc

    _check(dvoc.voc_eval, tmp_path, tag)


@pytest.mark.gpu
def test_device_matching_equals_oracle_random():
    from dafne_b200 import voc_eval as dvoc
    from tests.test_merge_nms import _rects

    rng = np.random.default_rng(11)
    nimg = 30
    gts, offs = [], [0]
    for _ in range(nimg):
        k = int(rng.integers(0, 40))
        gts.append(_rects(rng, k, 600, 20, 100, 3.0))
        offs.append(offs[-1] + k)
    gts = np.concatenate(gts, 0)
    nd = 3000
    det_image = rng.integers(0, nimg, nd).astype(np.int32)
    BB = _rects(rng, nd, 600, 20, 100, 3.0)
    for d in range(0, nd, 3):  # a third of the detections sit on a ground truth of their image
        lo, hi = offs[det_image[d]], offs[det_image[d] + 1]
        if hi > lo:
            BB[d] = gts[rng.integers(lo, hi)] + rng.normal(0, 2.0, 8)
    ov, jm = dvoc.match_detections(BB, det_image, gts, np.array(offs))
    for d in range(nd):
        want_ov, want_j = ovoc.match_detection(BB[d], gts[offs[det_image[d]]: offs[det_image[d] + 1]])
        assert ov[d] == want_ov and jm[d] == want_j
