"""Committed fixtures and the scripts that made them (run in the build container, where /root/reference exists)."""
import glob
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
STRIDES = [8, 16, 32, 64, 128]
REF_FIELDS = ("pred_corners", "pred_boxes", "scores", "centerness", "pred_classes", "fpn_levels", "locations")


def ref_postprocess_cases():
    """Names of the fixtures produced by the REFERENCE's own post-processing code (make_golden_postprocess_ref.py)."""
    return sorted(os.path.basename(f)[len("postprocess_ref_"):-4] for f in glob.glob(os.path.join(HERE, "postprocess_ref_*.npz")))


def load_ref_postprocess(name):
    g = np.load(os.path.join(HERE, f"postprocess_ref_{name}.npz"))
    C_, sort_c, twc, pre, post, dop = g["meta"].tolist()
    n_img = len(g["sizes"])
    return dict(
        logits=[g[f"logits{l}"] for l in range(5)], reg=[g[f"reg{l}"] for l in range(5)],
        ctr=[g[f"ctr{l}"] for l in range(5)], sizes=[tuple(r) for r in g["sizes"].tolist()],
        osz=[tuple(r) for r in g["osz"].tolist()], num_classes=C_, do_postprocess=bool(dop),
        kw=dict(pre_nms_topk=pre, post_nms_topk=post, sort_corners=bool(sort_c), thresh_with_ctr=bool(twc)),
        want=[{k: g[f"out{i}_{k}"] for k in REF_FIELDS} for i in range(n_img)])


def ulp_distance(a, b):
    """|a - b| in units in the last place, for positive float32 arrays."""
    return np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))


def assert_matches_reference_chain(got, want, what):
    """`got` (oracle or device result of one image) against what the reference's own code returned for it.

    Bit-identical: the number of detections, their order, classes, levels, locations, corner coordinates and boxes.
    scores / centerness: the reference evaluates torch's CPU float sigmoid, this repository the correctly rounded one
    (DESIGN.md "Ordering contract"); the two differ by at most 2 ulp per sigmoid, so every score may differ by at most
    SCORE_ULP and nothing else may: the rows that do differ are counted and the count is part of the message."""
    SCORE_ULP = 3  # sqrt(cls * ctr): two sigmoids <= 2 ulp each, halved by the square root, plus one rounding
    assert len(got["scores"]) == len(want["scores"]), (what, len(got["scores"]), len(want["scores"]))
    for k in ("pred_classes", "fpn_levels", "locations", "pred_corners", "pred_boxes"):
        assert np.array_equal(np.asarray(got[k]), want[k]), (what, k)
    for k, tol in (("scores", SCORE_ULP), ("centerness", 2)):
        d = ulp_distance(np.ascontiguousarray(got[k], np.float32), want[k])
        assert d.max(initial=0) <= tol, (what, k, int(d.max()), f"{int((d > 0).sum())} of {len(d)} rows differ")
    return int((ulp_distance(np.ascontiguousarray(got["scores"], np.float32), want["scores"]) > 0).sum())
