"""CPU tests that pin the oracle: against the reference's own artefacts (golden vectors generated from the reference's
sort_corners.py and polyiou.cpp, and the live files when /root/reference is present) and its known answers."""
import ctypes as C
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import postprocess as opost
from tests import golden

GOLD = os.path.join(os.path.dirname(__file__), "golden")
REF = "/root/reference"


def test_sort_quadrilateral_matches_reference_golden():
    g = np.load(os.path.join(GOLD, "sort_corners.npz"))
    got = opost.sort_quadrilateral(g["quads"])
    assert got.dtype == np.float32
    assert np.array_equal(got.view(np.uint32), g["sorted"].view(np.uint32))  # bit-exact incl. signed zeros


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree only exists in the build container")
def test_sort_quadrilateral_matches_reference_live():
    spec = importlib.util.spec_from_file_location("ref_sort_corners", os.path.join(REF, "dafne/utils/sort_corners.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(123)
    quads = rng.normal(0, 30, (3000, 8)).astype(np.float32)
    want = mod.sort_quadrilateral(torch.from_numpy(quads)).numpy()
    seq = mod.sort(torch.from_numpy(quads[:200])).numpy()  # the reference's sequential twin
    got = opost.sort_quadrilateral(quads)
    assert np.array_equal(got, want)
    assert np.array_equal(got[:200], seq)


def test_polyiou_f64_matches_reference_golden():
    g = np.load(os.path.join(GOLD, "polyiou_ref.npz"))
    got = opost.iou_poly_batch(g["p"], g["q"], double=True)
    assert np.array_equal(got, g["iou"])  # same algorithm, same precision, no contraction: identical doubles


def test_polyiou_known_answers():
    # polyiou.cpp:135-153 (its two main() functions) and the vectors SURVEY.md 8(c) derived from the vendored code
    sq = np.array([[0, 0, 1, 0, 1, 1, 0, 1]], np.float64)
    assert opost.iou_poly_batch(sq, sq + 0.5, True)[0] == pytest.approx(1 / 7, abs=1e-12)
    deg = np.array([[686, 2976, 709, 2976, 724, 2976, 701, 2976]], np.float64)
    assert opost.iou_poly_batch(deg, deg, True)[0] == 1.0  # union == 0 branch
    a = np.array([[0, 0, 2, 0, 2, 10, 0, 10]], np.float64)
    assert opost.iou_poly_batch(a, a + np.array([1, 0] * 4), True)[0] == pytest.approx(1 / 3, abs=1e-12)
    assert opost.iou_poly_batch(sq, sq.reshape(1, 4, 2)[:, ::-1].reshape(1, 8), True)[0] == pytest.approx(1.0)
    diamond = np.array([[0.5, -0.5, 1.5, 0.5, 0.5, 1.5, -0.5, 0.5]], np.float64)
    assert opost.iou_poly_batch(sq, diamond, True)[0] == pytest.approx(0.5, abs=1e-12)


@pytest.mark.skipif(opost.ref_lib() is None, reason="oracle/_ref not built")
def test_polyiou_f64_matches_reference_binary():
    ref = opost.ref_lib()
    rng = np.random.default_rng(5)
    p = rng.normal(0, 40, (4000, 8))
    q = p + rng.normal(0, 15, (4000, 8))
    out = np.empty(4000)
    dp = C.POINTER(C.c_double)
    ref.ref_iou_poly_batch(p.ctypes.data_as(dp), q.ctypes.data_as(dp), out.ctypes.data_as(dp), 4000)
    assert np.array_equal(opost.iou_poly_batch(p, q, True), out)


def test_polyiou_f32_close_to_f64_without_offset():
    g = np.load(os.path.join(GOLD, "polyiou_ref.npz"))
    p, q = g["p"][:2000].astype(np.float32), g["q"][:2000].astype(np.float32)
    f32 = opost.iou_poly_batch(p, q)
    f64 = opost.iou_poly_batch(p.astype(np.float64), q.astype(np.float64), True)
    assert np.abs(f32 - f64).max() < 5e-3  # coordinates < 300: fp32 is benign (offset 0)


def test_greedy_nms_basic_and_ties():
    boxes = np.array([[0, 0, 10, 0, 10, 10, 0, 10], [1, 1, 11, 1, 11, 11, 1, 11], [50, 50, 60, 50, 60, 60, 50, 60]],
                     np.float32)
    assert opost.greedy_nms(boxes, 0.1).tolist() == [0, 2]
    assert opost.greedy_nms(boxes, 0.9).tolist() == [0, 1, 2]
    assert opost.greedy_nms(boxes[:0], 0.1).tolist() == []


def _load_case(tag):
    g = np.load(os.path.join(GOLD, f"postprocess_{tag}.npz"))
    C_, sort_c, twc, pre, post = g["meta"].tolist()
    logits = [g[f"logits{l}"] for l in range(5)]
    reg = [g[f"reg{l}"] for l in range(5)]
    ctr = [g[f"ctr{l}"] for l in range(5)]
    sizes = [tuple(r) for r in g["sizes"].tolist()]
    osz = [tuple(r) for r in g["osz"].tolist()]
    return g, logits, reg, ctr, sizes, osz, dict(pre_nms_topk=pre, post_nms_topk=post, sort_corners=bool(sort_c),
                                                  thresh_with_ctr=bool(twc))


@pytest.mark.parametrize("tag", ["c15_sort", "c15_ctr", "c1_nosort"])
def test_postprocess_oracle_regression_and_invariants(tag):
    g, logits, reg, ctr, sizes, osz, kw = _load_case(tag)
    res = opost.postprocess(logits, reg, ctr, [8, 16, 32, 64, 128], sizes, osz, **kw)
    for i, r in enumerate(res):
        for k in ("pred_corners", "pred_boxes", "scores", "pred_classes", "canon", "locations"):
            assert np.array_equal(r[k], g[f"out{i}_{k}"]), (tag, i, k)
        s = r["scores"]
        assert np.all(s[:-1] >= s[1:])  # descending score
        assert np.all(s > 0.05)
        oh, ow = osz[i]
        b = r["pred_boxes"]
        assert np.all(b[:, 0] >= 0) and np.all(b[:, 2] <= ow) and np.all(b[:, 1] >= 0) and np.all(b[:, 3] <= oh)
        assert np.all(b[:, 2] > b[:, 0]) and np.all(b[:, 3] > b[:, 1])  # nonempty()


@pytest.mark.parametrize("name", golden.ref_postprocess_cases())
def test_postprocess_oracle_reproduces_the_reference_chain(name):
    """Fixtures made by EXECUTING the reference's dafne_outputs.py:733-925, nms/nms.py:10-92, sort_corners.py and
    one_stage_detector.py:45-98 (tests/golden/make_golden_postprocess_ref.py): threshold / top-k membership, decode,
    corner sort, class offsets, NMS order, the post-NMS cut with kthvalue, detectron2's scale / clip / nonempty and the
    do_postprocess gate -- same detections, same order, identical coordinates."""
    c = golden.load_ref_postprocess(name)
    res = opost.postprocess(c["logits"], c["reg"], c["ctr"], golden.STRIDES, c["sizes"], c["osz"],
                            do_postprocess=c["do_postprocess"], **c["kw"])
    assert len(res) == len(c["want"])
    assert max(len(w["scores"]) for w in c["want"]) > 60  # the fixture itself is non-trivial
    for i, (r, w) in enumerate(zip(res, c["want"])):
        golden.assert_matches_reference_chain(r, w, (name, i))
        assert r["image_size"] == c["osz"][i]


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("seed,C_,sort_c,twc", [(101, 15, True, True), (102, 2, False, False), (103, 15, True, False)])
def test_postprocess_oracle_reproduces_the_reference_chain_live(seed, C_, sort_c, twc):
    """The same comparison on fresh seeds, with the reference's code executed now (not only the stored fixtures)."""
    from tests.golden import make_golden_postprocess_ref as ref

    cfg = dict(C=C_, sort=sort_c, twc=twc, pre=300, post=120, seed=seed, bias=-2.2 if twc else (-3.2 if C_ > 2 else -1.6))
    for dop in (True, False):
        blob, want = ref.run_reference(cfg, dop)
        res = opost.postprocess([blob[f"logits{l}"] for l in range(5)], [blob[f"reg{l}"] for l in range(5)],
                                [blob[f"ctr{l}"] for l in range(5)], golden.STRIDES,
                                [tuple(r) for r in blob["sizes"].tolist()], [tuple(r) for r in blob["osz"].tolist()],
                                pre_nms_topk=300, post_nms_topk=120, sort_corners=sort_c, thresh_with_ctr=twc,
                                do_postprocess=dop)
        for i, (r, w) in enumerate(zip(res, want)):
            golden.assert_matches_reference_chain(r, w, (seed, dop, i))


def test_reference_chain_fixtures_exercise_every_cut():
    """The fixtures cover: per-level top-k binding (more candidates than pre_nms_topk), the post-NMS cut binding,
    rows dropped by the clipped-box filter, THRESH_WITH_CTR on and off, one class and sixteen, do_postprocess off."""
    names = golden.ref_postprocess_cases()
    assert {"c15_sort", "c15_ctr", "c16_nosort", "c1_sort", "c15_default_topk", "c15_sort_nopost"} <= set(names)
    c = golden.load_ref_postprocess("c15_sort")
    cls = opost.sigmoid_cr(c["logits"][0])
    assert int((cls[0] > 0.05).sum()) > c["kw"]["pre_nms_topk"]  # top-k binds at the first level
    assert len(c["want"][0]["scores"]) == c["kw"]["post_nms_topk"]  # the post-NMS cut binds
    nopost = golden.load_ref_postprocess("c15_sort_nopost")
    assert np.array_equal(nopost["want"][0]["pred_boxes"], c["want"][0]["pred_boxes"])  # boxes scaled either way
    assert not np.array_equal(nopost["want"][0]["pred_corners"], c["want"][0]["pred_corners"])


def test_postprocess_empty_image():
    logits = [np.full((1, 3, h, w), -20.0, np.float32) for h, w in ((8, 8), (4, 4), (2, 2), (1, 1), (1, 1))]
    reg = [np.zeros((1, 8, *t.shape[2:]), np.float32) for t in logits]
    ctr = [np.zeros((1, 1, *t.shape[2:]), np.float32) for t in logits]
    res = opost.postprocess(logits, reg, ctr, [8, 16, 32, 64, 128], [(64, 64)])
    assert len(res) == 1 and len(res[0]["scores"]) == 0 and res[0]["pred_corners"].shape == (0, 8)


def test_sigmoid_cr_within_2ulp_of_torch():
    x = torch.linspace(-12, 6, 20001)
    a = opost.sigmoid_cr(x.numpy())
    b = torch.sigmoid(x).numpy()
    ulp = np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))
    assert ulp.max() <= 2
