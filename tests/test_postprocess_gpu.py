"""GPU parity of the post-processing kernels against the oracle: bit-exact (integer / index work and the faithful fp32
arithmetic), through the C ABI, on the committed golden fixtures and on seeded random inputs."""
import os

import numpy as np
import pytest
import torch

from oracle import postprocess as opost
from tests import golden

GOLD = os.path.join(os.path.dirname(__file__), "golden")
pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda:0")


def test_sort_quadrilateral_matches_reference_golden_on_gpu():
    from dafne_b200.modeling import sort_quadrilateral

    g = np.load(os.path.join(GOLD, "sort_corners.npz"))
    got = sort_quadrilateral(torch.from_numpy(g["quads"]).to(_dev())).cpu().numpy()
    assert np.array_equal(got.view(np.uint32), g["sorted"].view(np.uint32))


def test_sort_quadrilateral_empty():
    from dafne_b200.modeling import sort_quadrilateral

    assert sort_quadrilateral(torch.zeros(0, 8, device=_dev())).shape == (0, 8)


@pytest.mark.parametrize("offset", [0.0, 4400.0, 15400.0])
def test_poly_iou_bit_exact_vs_oracle_f32(offset):
    """Including the class-offset regimes where fp32 cancels catastrophically: the kernel must
    reproduce the oracle's fp32 result bit for bit, whatever its distance from the true IoU."""
    from dafne_b200.modeling import poly_iou

    g = np.load(os.path.join(GOLD, "polyiou_ref.npz"))
    p = (g["p"] + offset).astype(np.float32)
    q = (g["q"] + offset).astype(np.float32)
    want = opost.iou_poly_batch(p, q)
    got = poly_iou(torch.from_numpy(p).to(_dev()), torch.from_numpy(q).to(_dev())).cpu().numpy()
    same = got.view(np.uint32) == want.view(np.uint32)
    assert same.all(), f"{(~same).sum()} of {len(same)} differ, first at {np.nonzero(~same)[0][:5]}"
    if offset == 0.0:
        assert np.abs(got.astype(np.float64) - g["iou"]).max() < 5e-3  # and close to the reference's double result


def _pair_filter(p, q):
    import ctypes as C

    from dafne_b200 import _capi

    fired = torch.zeros(p.shape[0], dtype=torch.uint8, device=p.device)
    _capi.check(_capi.lib().dafne_poly_pair_filter(p.data_ptr(), q.data_ptr(), fired.data_ptr(), p.shape[0],
                                                   C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                "dafne_poly_pair_filter")
    return fired.bool()


def _rot_rects_t(gen, n, wmin, wmax, aspect, cx, cy, dev):
    """n rotated rectangles [n, 8] on the device: width U(wmin, wmax), height width/aspect, centre given."""
    w = torch.rand(n, generator=gen, device=dev) * (wmax - wmin) + wmin
    h = w / aspect
    a = torch.rand(n, generator=gen, device=dev) * 3.14159265
    ca, sa = torch.cos(a), torch.sin(a)
    dx = torch.stack([-w, w, w, -w], 1) * 0.5
    dy = torch.stack([-h, -h, h, h], 1) * 0.5
    x = cx[:, None] + dx * ca[:, None] - dy * sa[:, None]
    y = cy[:, None] + dx * sa[:, None] + dy * ca[:, None]
    return torch.stack([x, y], 2).reshape(n, 8).float().contiguous()


REGIMES = ["same_class", "cross_class", "near_margin", "angular_touch", "integer_grid", "tiny_boxes", "huge_boxes",
           "near_origin", "degenerate"]


def _regime_pairs(regime, n=1 << 22):
    """(p, q): n pairs of quadrilaterals [n, 8] on the device for one regime of the NMS filter tests."""
    dev = _dev()
    gen = torch.Generator(device=dev).manual_seed(sum(map(ord, regime)))
    U = lambda lo, hi: torch.rand(n, generator=gen, device=dev) * (hi - lo) + lo  # noqa: E731
    span = 1100.0
    if regime == "same_class":
        off = torch.randint(0, 15, (n,), generator=gen, device=dev).float() * span
        p = _rot_rects_t(gen, n, 8, 120, 3.0, U(0, 1024) + off, U(0, 1024) + off, dev)
        q = _rot_rects_t(gen, n, 8, 120, 3.0, U(0, 1024) + off, U(0, 1024) + off, dev)
    elif regime == "cross_class":
        o1 = torch.randint(0, 15, (n,), generator=gen, device=dev).float() * span
        o2 = torch.randint(0, 15, (n,), generator=gen, device=dev).float() * span
        p = _rot_rects_t(gen, n, 8, 400, 3.0, U(0, 1024) + o1, U(0, 1024) + o1, dev)
        q = _rot_rects_t(gen, n, 8, 400, 3.0, U(0, 1024) + o2, U(0, 1024) + o2, dev)
    elif regime == "near_margin":  # q a few box sizes away from p, every direction
        off = torch.randint(1, 15, (n,), generator=gen, device=dev).float() * span
        cx, cy = U(0, 1024) + off, U(0, 1024) + off
        p = _rot_rects_t(gen, n, 20, 60, 3.0, cx, cy, dev)
        ang, dist = U(0, 6.2832), U(0, 250)
        q = _rot_rects_t(gen, n, 20, 60, 3.0, cx + dist * torch.cos(ang), cy + dist * torch.sin(ang), dev)
    elif regime == "angular_touch":  # q next to p's cone seen from the origin (tiny angular gaps either side), any radius
        off = torch.randint(0, 15, (n,), generator=gen, device=dev).float() * span
        cx, cy = U(0, 1024) + off, U(0, 1024) + off
        p = _rot_rects_t(gen, n, 20, 100, 3.0, cx, cy, dev)
        r = torch.sqrt(cx * cx + cy * cy)
        tang, rad = U(-160, 160), U(-900, 900)  # tangential shift of about one box, radial shift across classes
        scale = (r + rad).clamp(min=40.0) / r
        qx, qy = cx * scale - cy / r * tang, cy * scale + cx / r * tang
        q = _rot_rects_t(gen, n, 20, 100, 3.0, qx.clamp(min=60.0), qy.clamp(min=60.0), dev)
    elif regime == "integer_grid":  # axis-aligned integer boxes: exactly collinear edges, exact zeros in the clips
        off = torch.randint(0, 15, (n,), generator=gen, device=dev).float() * 1024.0

        def grid_boxes():
            x0 = torch.randint(1, 200, (n,), generator=gen, device=dev).float() + off
            y0 = torch.randint(1, 200, (n,), generator=gen, device=dev).float() + off
            w = torch.randint(1, 40, (n,), generator=gen, device=dev).float()
            h = torch.randint(1, 40, (n,), generator=gen, device=dev).float()
            return torch.stack([x0, y0, x0 + w, y0, x0 + w, y0 + h, x0, y0 + h], 1).contiguous()

        p, q = grid_boxes(), grid_boxes()
    elif regime == "tiny_boxes":
        off = torch.randint(0, 15, (n,), generator=gen, device=dev).float() * span
        p = _rot_rects_t(gen, n, 0.5, 6, 2.0, U(0, 300) + off, U(0, 300) + off, dev)
        q = _rot_rects_t(gen, n, 0.5, 6, 2.0, U(0, 300) + off, U(0, 300) + off, dev)
    elif regime == "huge_boxes":
        off = torch.randint(0, 15, (n,), generator=gen, device=dev).float() * 2600.0
        p = _rot_rects_t(gen, n, 200, 900, 3.0, U(0, 1024) + off, U(0, 1024) + off, dev)
        q = _rot_rects_t(gen, n, 10, 900, 8.0, U(0, 1024) + off, U(0, 1024) + off, dev)
    elif regime == "near_origin":  # class 0: coordinates around and below 1, negative, mixed with far boxes
        p = _rot_rects_t(gen, n, 2, 80, 3.0, U(-20, 200), U(-20, 200), dev)
        q = _rot_rects_t(gen, n, 2, 80, 3.0, U(-20, 2000), U(-20, 2000), dev)
    else:  # degenerate: zero-area quads, repeated vertices, edges through the origin direction
        off = torch.randint(0, 15, (n,), generator=gen, device=dev).float() * span
        p = _rot_rects_t(gen, n, 8, 120, 3.0, U(0, 1024) + off, U(0, 1024) + off, dev)
        q = _rot_rects_t(gen, n, 8, 120, 3.0, U(0, 1024) + off, U(0, 1024) + off, dev)
        p[::3, 4:8] = p[::3, 0:4]                       # collapsed to a segment
        q[1::3] = q[1::3, :2].repeat(1, 4)              # a single point
        k = torch.arange(2, n, 3, device=dev)           # a radial sliver: two vertices on one ray from the origin
        q[k, 2:4] = q[k, 0:2] * 1.25
    return p, q


@pytest.mark.parametrize("regime", REGIMES)
def test_nms_prefilter_never_skips_a_nonzero_iou(regime):
    """Contract of polyiou.cuh::pair_inter_is_zero: wherever the NMS skips the polygon clip, the faithful fp32
    arithmetic (dafne_poly_iou == the oracle's float instantiation, bit for bit) yields EXACTLY 0 -- including the
    class-offset regimes where fp32 IoU of disjoint boxes is noise. 4M pairs per regime."""
    from dafne_b200.modeling import poly_iou

    p, q = _regime_pairs(regime)
    fired = _pair_filter(p, q)
    iou = poly_iou(p, q)
    bad = fired & (iou != 0)
    assert int(bad.sum()) == 0, f"{regime}: filter fired on {int(bad.sum())} pairs with IoU != 0, e.g. {iou[bad][:4]}"
    if regime in ("same_class", "cross_class", "tiny_boxes"):
        assert fired.float().mean() > 0.3, f"{regime}: the filter should fire on most separated pairs"
    if regime == "near_origin":
        assert int((fired & (p.min(1).values < 1.0)).sum()) == 0  # boxes with a coordinate < 1 are never eligible


def _term_filter(p, q):
    import ctypes as C

    from dafne_b200 import _capi

    fired = torch.zeros(p.shape[0], dtype=torch.int16, device=p.device)
    nonzero = torch.zeros(p.shape[0], dtype=torch.int16, device=p.device)
    _capi.check(_capi.lib().dafne_poly_term_filter(p.data_ptr(), q.data_ptr(), fired.data_ptr(), nonzero.data_ptr(),
                                                   p.shape[0], C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                "dafne_poly_term_filter")
    return fired.to(torch.int32) & 0xFFFF, nonzero.to(torch.int32) & 0xFFFF


@pytest.mark.parametrize("regime", REGIMES)
def test_nms_term_filter_never_skips_a_nonzero_term(regime):
    """Contract of polyiou.cuh::term_is_zero, the per-term form of the filter: a signed triangle overlap the NMS does
    not evaluate is EXACTLY zero in the faithful fp32 arithmetic (the skipped terms enter the sum as +0). 2M pairs =
    32M terms per regime, the same regimes as the pair filter (class offsets, near margins, degenerate quads)."""
    p, q = _regime_pairs(regime, 1 << 21)
    fired, nonzero = _term_filter(p, q)
    bad = (fired & nonzero) != 0
    assert int(bad.sum()) == 0, f"{regime}: {int(bad.sum())} pairs have a skipped term that is not zero"
    live = 16 - torch.tensor([bin(v).count("1") for v in range(1 << 16)], device=p.device)[fired.long()]
    if regime in ("same_class", "near_margin", "angular_touch"):
        # the point of the filter: of the pairs whose wedges overlap, about half of the 16 terms are provably zero
        assert live.float().mean() < 12.5, f"{regime}: {live.float().mean():.2f} of 16 terms left to clip"


def _random_boxes(rng, n, span=400, ncls=15, small=False):
    from tests.golden.make_golden import rot_rects

    w, h = (12, 4) if small else (48, 16)
    boxes = rot_rects(rng, n, w, h, center=(0, span), jitter=1.0)
    scores = rng.uniform(0.05, 1.0, n).astype(np.float32)
    classes = rng.integers(0, ncls, n).astype(np.int64)
    return boxes, scores, classes


def _oracle_nms(boxes, scores, classes, thr, merge=True):
    order = np.lexsort((np.arange(len(scores)), -scores.astype(np.float64)))
    idx = classes.copy()
    if merge:
        idx[idx == 5] = 4
    span = (boxes.max() - boxes.min()) + np.float32(1.0)
    shifted = boxes + (idx.astype(np.float32) * span)[:, None]
    keep = opost.greedy_nms(shifted[order], thr)
    return order[keep]


@pytest.mark.parametrize("n,span,ncls,small", [(1, 100, 1, False), (63, 200, 3, False), (64, 200, 15, False),
                                              (65, 100, 1, False), (700, 300, 15, False), (2000, 500, 15, True),
                                              (3000, 1024, 1, False)])
def test_poly_nms_matches_oracle(n, span, ncls, small):
    from dafne_b200.modeling import batched_nms_poly

    rng = np.random.default_rng(n)
    boxes, scores, classes = _random_boxes(rng, n, span, ncls, small)
    want = _oracle_nms(boxes, scores, classes, 0.1)
    got = batched_nms_poly(torch.from_numpy(boxes).to(_dev()), torch.from_numpy(scores).to(_dev()),
                           torch.from_numpy(classes).to(_dev()), 0.1).cpu().numpy()
    assert np.array_equal(got, want), f"kept {len(got)} vs {len(want)}"


def test_poly_nms_score_ties_and_duplicates():
    """Exact score ties: order is defined as ascending input index; duplicates of a box are suppressed by the first."""
    from dafne_b200.modeling import batched_nms_poly

    rng = np.random.default_rng(0)
    boxes, scores, classes = _random_boxes(rng, 300, 150, 2)
    boxes[100:200] = boxes[0:100]
    scores[:] = np.repeat(rng.uniform(0.1, 0.9, 30).astype(np.float32), 10)
    classes[100:200] = classes[0:100]
    want = _oracle_nms(boxes, scores, classes, 0.1)
    got = batched_nms_poly(torch.from_numpy(boxes).to(_dev()), torch.from_numpy(scores).to(_dev()),
                           torch.from_numpy(classes).to(_dev()), 0.1).cpu().numpy()
    assert np.array_equal(got, want)


def test_poly_gpu_nms_host_dropin():
    """The reference's FFI shape: dets [n, 9] host numpy (offsets applied by the caller) -> list of kept indices."""
    from dafne_b200.modeling import poly_gpu_nms

    rng = np.random.default_rng(11)
    boxes, scores, _ = _random_boxes(rng, 500, 200, 1)
    dets = np.hstack([boxes, scores[:, None]]).astype(np.float32)
    order = np.lexsort((np.arange(500), -scores.astype(np.float64)))
    want = order[opost.greedy_nms(boxes[order], 0.1)]
    assert poly_gpu_nms(dets, 0.1, 0) == want.tolist()
    assert poly_gpu_nms(dets[:0], 0.1, 0) == []


def test_vehicle_merge_hack():
    from dafne_b200.modeling import batched_nms_poly

    b = np.array([[0, 0, 10, 0, 10, 10, 0, 10], [1, 1, 11, 1, 11, 11, 1, 11]], np.float32)
    s = torch.tensor([0.9, 0.8], device=_dev())
    bt = torch.from_numpy(b).to(_dev())
    assert batched_nms_poly(bt, s, torch.tensor([4, 5], device=_dev()), 0.1).tolist() == [0]  # 5 is treated as 4
    assert batched_nms_poly(bt, s, torch.tensor([3, 5], device=_dev()), 0.1).tolist() == [0, 1]
    assert batched_nms_poly(bt, s, torch.tensor([4, 5], device=_dev()), 0.1, vehicle_merge=False).tolist() == [0, 1]


def _engine(C_, sort_c, twc, pre, post):
    from dafne_b200.engine import DafneEngine
    from dafne_b200.spec import ModelSpec

    spec = ModelSpec(resnet_depth=50, num_classes=C_, sort_corners=sort_c, thresh_with_ctr=twc, pre_nms_topk=pre,
                     post_nms_topk=post)
    return DafneEngine(spec, _dev()), spec


def _compare(res, dets, counts, exact=True):
    dets = dets.cpu().numpy()
    counts = counts.cpu().numpy()
    for i, r in enumerate(res):
        n = len(r["scores"])
        assert counts[i] == n, (i, counts[i], n)
        g = dets[i, :n]
        assert np.array_equal(g[:, 18].view(np.uint32).astype(np.int64), r["canon"])  # candidate indices: bit-exact
        assert np.array_equal(g[:, 14].astype(np.int64), r["pred_classes"])
        assert np.array_equal(g[:, 15].astype(np.int64), r["fpn_levels"])
        assert np.array_equal(g[:, 0:8], r["pred_corners"])
        assert np.array_equal(g[:, 8:12], r["pred_boxes"])
        assert np.array_equal(g[:, 12], r["scores"])
        assert np.array_equal(g[:, 13], r["centerness"])
        assert np.array_equal(g[:, 16:18], r["locations"])


@pytest.mark.parametrize("tag", ["c15_sort", "c15_ctr", "c1_nosort"])
def test_postprocess_external_matches_golden(tag):
    g = np.load(os.path.join(GOLD, f"postprocess_{tag}.npz"))
    C_, sort_c, twc, pre, post = g["meta"].tolist()
    eng, spec = _engine(C_, bool(sort_c), bool(twc), pre, post)
    logits = [torch.from_numpy(g[f"logits{l}"]) for l in range(5)]
    reg = [torch.from_numpy(g[f"reg{l}"]) for l in range(5)]
    ctr = [torch.from_numpy(g[f"ctr{l}"]) for l in range(5)]
    sizes = [tuple(r) for r in g["sizes"].tolist()]
    osz = [tuple(r) for r in g["osz"].tolist()]
    dets, counts = eng.postprocess_external(logits, reg, ctr, sizes, osz, True)
    res = [{k[len(f"out{i}_"):]: g[k] for k in g.files if k.startswith(f"out{i}_")} for i in range(2)]
    _compare(res, dets, counts)
    eng.close()


@pytest.mark.parametrize("name", golden.ref_postprocess_cases())
def test_postprocess_device_reproduces_the_reference_chain(name):
    """The device post-processing against fixtures made by EXECUTING the reference's own dafne_outputs.py:733-925,
    nms/nms.py:10-92, sort_corners.py and one_stage_detector.py:45-98 (tests/golden/make_golden_postprocess_ref.py):
    identical detections in identical order, identical classes / levels / coordinates / boxes; scores within the
    sigmoid convention's ulps (tests/golden/__init__.py::assert_matches_reference_chain)."""
    c = golden.load_ref_postprocess(name)
    kw = c["kw"]
    eng, spec = _engine(c["num_classes"], kw["sort_corners"], kw["thresh_with_ctr"], kw["pre_nms_topk"],
                        kw["post_nms_topk"])
    dets, counts = eng.postprocess_external([torch.from_numpy(t) for t in c["logits"]],
                                            [torch.from_numpy(t) for t in c["reg"]],
                                            [torch.from_numpy(t) for t in c["ctr"]], c["sizes"], c["osz"],
                                            c["do_postprocess"])
    dets, counts = dets.cpu().numpy(), counts.cpu().numpy()
    for i, w in enumerate(c["want"]):
        n = int(counts[i])
        g = dets[i, :n]
        got = dict(pred_corners=g[:, 0:8], pred_boxes=g[:, 8:12], scores=g[:, 12], centerness=g[:, 13],
                   pred_classes=g[:, 14].astype(np.int64), fpn_levels=g[:, 15].astype(np.int64), locations=g[:, 16:18])
        golden.assert_matches_reference_chain(got, w, (name, i))
    eng.close()


@pytest.mark.parametrize("C_,sort_c,twc,seed,scale", [(15, True, False, 21, 1.0), (15, True, True, 22, 1.0),
                                                      (1, True, False, 23, 2.0), (16, False, False, 24, 1.0),
                                                      (2, False, True, 25, 0.5)])
def test_postprocess_random_heads_match_oracle(C_, sort_c, twc, seed, scale):
    """Per-level top-k cap active at P3, ragged image sizes, rescaling to a different output size, odd level sizes."""
    rng = np.random.default_rng(seed)
    N = 3
    hw = [(50, 38), (25, 19), (13, 10), (7, 5), (4, 3)]
    strides = [8, 16, 32, 64, 128]
    bias = -5.6 if twc else -2.6
    logits = [rng.normal(bias, 1.3, (N, C_, h, w)).astype(np.float32) for h, w in hw]
    base = np.array([-3, -1, 3, -1, 3, 1, -3, 1], np.float32).reshape(1, 8, 1, 1) * scale
    reg = [(base + rng.normal(0, 0.8, (N, 8, h, w))).astype(np.float32) for h, w in hw]
    ctr = [rng.normal(0, 1, (N, 1, h, w)).astype(np.float32) for h, w in hw]
    sizes = [(400, 304), (380, 290), (333, 257)]
    osz = [(800, 608), (380, 290), (100, 80)]
    eng, spec = _engine(C_, sort_c, twc, 300, 200)
    res = opost.postprocess(logits, reg, ctr, strides, sizes, osz, pre_nms_topk=300, post_nms_topk=200,
                            sort_corners=sort_c, thresh_with_ctr=twc)
    dets, counts = eng.postprocess_external([torch.from_numpy(t) for t in logits], [torch.from_numpy(t) for t in reg],
                                            [torch.from_numpy(t) for t in ctr], sizes, osz, True)
    assert max(len(r["scores"]) for r in res) > 20  # sanity of the test inputs themselves
    _compare(res, dets, counts)
    # do_postprocess=False (the TTA call shape, tta.py:190-194): boxes scaled / clipped / filtered, corners untouched
    res2 = opost.postprocess(logits, reg, ctr, strides, sizes, osz, pre_nms_topk=300, post_nms_topk=200,
                             sort_corners=sort_c, thresh_with_ctr=twc, do_postprocess=False)
    dets2, counts2 = eng.postprocess_external([torch.from_numpy(t) for t in logits], [torch.from_numpy(t) for t in reg],
                                              [torch.from_numpy(t) for t in ctr], sizes, osz, False)
    _compare(res2, dets2, counts2)
    eng.close()


def test_postprocess_empty_and_all_pass():
    eng, spec = _engine(3, True, False, 50, 20)
    hw = [(8, 8), (4, 4), (2, 2), (1, 1), (1, 1)]
    strides = [8, 16, 32, 64, 128]
    rng = np.random.default_rng(9)
    # image 0: nothing above threshold; image 1: every (location, class) passes -> top-k everywhere
    logits = [np.stack([np.full((3, h, w), -20.0), rng.normal(3.0, 1.0, (3, h, w))]).astype(np.float32) for h, w in hw]
    reg = [rng.normal(0, 2, (2, 8, h, w)).astype(np.float32) for h, w in hw]
    ctr = [rng.normal(0, 1, (2, 1, h, w)).astype(np.float32) for h, w in hw]
    sizes = [(64, 64), (64, 64)]
    res = opost.postprocess(logits, reg, ctr, strides, sizes, None, pre_nms_topk=50, post_nms_topk=20)
    dets, counts = eng.postprocess_external([torch.from_numpy(t) for t in logits], [torch.from_numpy(t) for t in reg],
                                            [torch.from_numpy(t) for t in ctr], sizes, None, True)
    assert counts.tolist()[0] == 0 and len(res[0]["scores"]) == 0
    _compare(res, dets, counts)
    eng.close()


def test_post_nms_cut_keeps_every_score_tied_with_the_kth():
    """dafne_outputs.py:916-923: `kthvalue` + `scores >= thr` keeps ALL detections whose score equals the k-th best, so
    an image can return more than POST_NMS_TOPK_TEST rows. Exact ties at the cut, on the device: 10 distinct scores,
    then 8 boxes with identical logits and centerness (identical scores), then lower ones; post_nms_topk = 12 lands
    inside the tie group -> 18 rows. Boxes are 4 px wide on a stride-8 grid, so the NMS suppresses nothing."""
    C_, H, W = 1, 12, 12
    eng, spec = _engine(C_, False, False, 2000, 12)
    hw = [(H, W), (6, 6), (3, 3), (2, 2), (1, 1)]
    logits = [np.full((1, C_, h, w), -20.0, np.float32) for h, w in hw]
    ctr = [np.zeros((1, 1, h, w), np.float32) for h, w in hw]
    base = np.array([-0.25, -0.25, 0.25, -0.25, 0.25, 0.25, -0.25, 0.25], np.float32).reshape(1, 8, 1, 1)  # stride units
    reg = [np.tile(base, (1, 1, h, w)).astype(np.float32) for h, w in hw]
    flat = logits[0].reshape(-1)
    flat[0:10] = np.linspace(3.0, 2.1, 10, dtype=np.float32)  # 10 distinct, best first
    flat[20:28] = 1.5                                          # 8 exact ties
    flat[40:60] = np.linspace(1.0, 0.2, 20, dtype=np.float32)  # 20 lower ones
    sizes = [(H * 8, W * 8)]
    res = opost.postprocess(logits, reg, ctr, [8, 16, 32, 64, 128], sizes, None, pre_nms_topk=2000, post_nms_topk=12,
                            sort_corners=False, thresh_with_ctr=False)
    assert len(res[0]["scores"]) == 18 and len(np.unique(res[0]["scores"][10:])) == 1
    t = [[torch.from_numpy(a) for a in x] for x in (logits, reg, ctr)]
    dets, counts = eng.postprocess_external(t[0], t[1], t[2], sizes, None, True)
    _compare(res, dets, counts)
    assert int(counts[0]) == 18
    # the reference's semantics when the ties do not fit the caller's buffer: the count still says 18, rows are cut
    dets2, counts2 = eng.postprocess_external(t[0], t[1], t[2], sizes, None, True, capacity=14)
    assert int(counts2[0]) == 18 and dets2.shape[1] == 14
    assert np.array_equal(dets2[0].cpu().numpy()[:, 12], res[0]["scores"][:14])
    eng.close()


def test_rows_past_the_count_are_zero_and_buffers_are_reusable():
    """The finalize kernel writes every element of the record: a caller-owned DetectionWire reused across batches never
    shows rows of an earlier batch (no per-step clear on the host side)."""
    from dafne_b200.engine import DetectionWire

    c = golden.load_ref_postprocess("c15_sort")
    kw = c["kw"]
    eng, spec = _engine(c["num_classes"], kw["sort_corners"], kw["thresh_with_ctr"], kw["pre_nms_topk"],
                        kw["post_nms_topk"])
    t = [[torch.from_numpy(a) for a in c[k]] for k in ("logits", "reg", "ctr")]
    dets, counts = eng.postprocess_external(t[0], t[1], t[2], c["sizes"], c["osz"], True)
    for i, n in enumerate(counts.tolist()):
        assert n > 0 and float(dets[i, n:].abs().max()) == 0.0
    w = DetectionWire(2, 8, _dev())
    assert w.dets.shape == (2, 8, 20) and w.counts.shape == (2,) and w.counts.dtype == torch.int32
    assert w.dets.data_ptr() == w.buf.data_ptr() and w.counts.data_ptr() == w.buf.data_ptr() + 2 * 8 * 20 * 4
    eng.close()
