"""Patch-merge polygon NMS: oracle vs golden vectors produced by the reference's own
py_cpu_nms_poly_fast (CPU), device vs both (GPU)."""
import os

import numpy as np
import pytest

from oracle import merge_nms as omerge

GOLD = os.path.join(os.path.dirname(__file__), "golden", "patch_merge_nms.npz")
CASES = ["one", "sparse", "ships", "dense", "dup"]


def _rects(rng, n, extent, wmin, wmax, aspect):
    cx, cy = rng.uniform(0, extent, n), rng.uniform(0, extent, n)
    w = rng.uniform(wmin, wmax, n)
    h = w / aspect
    a = rng.uniform(0, np.pi, n)
    dx = np.stack([-w, w, w, -w], 1) / 2
    dy = np.stack([-h, -h, h, h], 1) / 2
    x = cx[:, None] + dx * np.cos(a)[:, None] - dy * np.sin(a)[:, None]
    y = cy[:, None] + dx * np.sin(a)[:, None] + dy * np.cos(a)[:, None]
    return np.stack([x, y], 2).reshape(n, 8)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    g = np.load(GOLD)
    for thr in (0.1, 0.3):
        assert omerge.py_cpu_nms_poly_fast(g[f"{name}_dets"], thr) == g[f"{name}_keep_{thr}"].tolist()


def test_oracle_poly2origpoly_and_dict():
    assert omerge.poly2origpoly([10, 20, 30, 40], 100, 200, "0.5") == [220.0, 440.0, 260.0, 480.0]
    g = np.load(GOLD)
    d = {"P1": g["dup_dets"].tolist(), "P2": g["one_dets"].tolist()}
    out = omerge.nmsbynamedict(d, 0.1)
    assert [len(out["P1"]), len(out["P2"])] == [len(g["dup_keep_0.1"]), 1]


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_device_matches_reference_golden(name):
    from dafne_b200 import merge

    g = np.load(GOLD)
    for thr in (0.1, 0.3):
        assert merge.py_cpu_nms_poly_fast(g[f"{name}_dets"], thr) == g[f"{name}_keep_{thr}"].tolist()


@pytest.mark.gpu
def test_device_batch_matches_oracle_random():
    """Many lists in one call, sizes around the 64-box block boundaries, DOTA-scale coordinates, heavy overlaps."""
    from dafne_b200 import merge

    rng = np.random.default_rng(3)
    lists = []
    for n, extent in ((0, 10), (1, 50), (63, 300), (64, 300), (65, 300), (129, 500), (1000, 1200), (2500, 3000)):
        boxes = _rects(rng, n, extent, 15, 120, rng.uniform(1.5, 7.0))
        scores = rng.permutation(n) / max(n, 1) * 0.9 + 0.05 + rng.uniform(0, 1e-5, n)
        lists.append(np.concatenate([boxes, scores[:, None]], 1))
    got = merge.nms_many(lists, 0.1)
    for dets, k in zip(lists, got):
        assert k == omerge.py_cpu_nms_poly_fast(dets, 0.1)


@pytest.mark.gpu
def test_device_ties_duplicates_and_degenerate():
    from dafne_b200 import merge

    rng = np.random.default_rng(5)
    boxes = _rects(rng, 40, 200, 20, 60, 3.0)
    dets = np.concatenate([boxes, np.full((40, 1), 0.5)], 1)  # all scores equal: ties by ascending index
    dets[10:20, :8] = dets[0:10, :8]  # exact duplicates: the later copy is dropped
    dets[30, :8] = np.tile(dets[30, :2], 4)  # a point: zero-area polygon, "+1" hbox area
    assert merge.py_cpu_nms_poly_fast(dets, 0.1) == omerge.py_cpu_nms_poly_fast(dets, 0.1)
    assert merge.py_cpu_nms_poly_fast(np.zeros((0, 9)), 0.1) == []


@pytest.mark.gpu
def test_merge_lines_like_mergesingle():
    from dafne_b200 import merge

    rng = np.random.default_rng(9)
    lines, dets_by_img = [], {}
    for img in ("P0006", "P0011"):
        for (ox, oy) in ((0, 0), (824, 0), (0, 824)):
            b = np.round(_rects(rng, 30, 1024, 30, 120, 4.0), 1)
            for k in range(30):
                conf = round(float(rng.uniform(0.05, 1.0)), 6)
                lines.append(f"{img}__1__{ox}___{oy} {conf} " + " ".join(str(v) for v in b[k]))
                dets_by_img.setdefault(img, []).append(omerge.poly2origpoly(list(b[k]), ox, oy, "1") + [conf])
    out = merge.merge_lines(lines)
    want = omerge.nmsbynamedict(dets_by_img, 0.1)
    assert len(out) == sum(len(v) for v in want.values())
    for line in out:
        parts = line.split(" ")
        assert [float(v) for v in parts[2:]] + [float(parts[1])] in want[parts[0]]
