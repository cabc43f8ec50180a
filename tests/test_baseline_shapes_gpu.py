"""GPU parity at the BASELINE.json shapes (the sizes the bench is quoted on), through the C ABI:

  r50_1024    configs[1]: dota-1.0 1024.yaml flavour (R50, C = 15, SORT_CORNERS), 1 x 1024 x 1024
  r101_1024   configs[2]: dota-1.0_r101_ms (R101, C = 15, SORT_CORNERS + THRESH_WITH_CTR), 1 x 1024 x 1024
  hrsc_mixed  configs[4]: hrsc_r50_ms (R50, C = 1), one 512^2 + one 800^2 + one 1024^2 image zero-padded into one batch
              like ImageList.from_tensors does (one_stage_detector.py:100-107)

At 1024^2 the launch plan takes the branches the toy-size tests never reach: 16 x 8 pixel tiles, multi-wave grouped
launches over the 128^2 ... 8^2 levels, deep rings, the residual 1x1 convolutions on 128-wide tiles. Per case:
  * every named activation vs the quantisation-matched oracle (o16): rel L2 <= 8e-3;
  * head outputs vs the reference's fp32 arithmetic: |d logit|, |d ctr| <= 3e-2, |d reg| <= 5e-2 (stride units);
  * post-processing of the GPU's own head outputs vs the oracle: BIT-EXACT indices, classes, coordinates, scores;
  * end to end vs the fp32 oracle: the MEASURED agreement (share of the reference's detections reproduced, max |d score|,
    max |d coord| over the matched ones) is asserted against a floor and written to gpurun_out/parity_baseline_shapes.json
    (DESIGN.md section 2 quotes it) -- for kernel-vs-o16, kernel-vs-fp32 and o16-vs-fp32.
The CPU oracle needs about 1-2 s per image and precision on the box's host cores, so every case is one small batch.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import model as omodel
from oracle import postprocess as opost

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = {
    "r50_1024": dict(cfg="configs/dota10_r50_1024.yaml", sizes=[(1024, 1024)], seed=1234, synth={}),
    "r101_1024": dict(cfg="configs/dota10_r101_ms.yaml", sizes=[(1024, 1024)], seed=1235, synth={}),
    "hrsc_mixed": dict(cfg="configs/hrsc_r50_ms.yaml", sizes=[(512, 512), (800, 800), (1024, 1024)], seed=1236,
                       synth=dict(cls_bias=-2.9, base_quad=(-4.0, -0.5, 4.0, -0.5, 4.0, 0.5, -4.0, 0.5))),
}
_REPORT = {}


def _record(case_name, key, value):
    """Measured numbers go to gpurun_out/parity_baseline_shapes.json (rewritten after every update)."""
    _REPORT.setdefault(case_name, {})[key] = value
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_baseline_shapes.json"), "w") as f:
        json.dump(_REPORT, f, indent=1, sort_keys=True)


def _spec(cfg_file):
    from dafne_b200.config import get_cfg
    from dafne_b200.spec import ModelSpec

    cfg = get_cfg()
    cfg.merge_from_file(os.path.join(ROOT, cfg_file))
    return ModelSpec.from_cfg(cfg)


@pytest.fixture(scope="module", params=list(CASES))
def case(request):
    from dafne_b200.engine import DafneEngine
    from dafne_b200.weights import synthetic_state_dict

    c = CASES[request.param]
    spec = _spec(c["cfg"])
    sd = synthetic_state_dict(spec, 0, **c["synth"])
    eng = DafneEngine(spec, torch.device("cuda:0"))
    eng.load_state_dict(sd)
    eng.keep_activations(True)
    g = torch.Generator().manual_seed(c["seed"])
    imgs = [torch.randint(0, 256, (3, h, w), dtype=torch.uint8, generator=g) for h, w in c["sizes"]]
    H = max(h for h, _ in c["sizes"])
    W = max(w for _, w in c["sizes"])
    batch_u8 = torch.zeros(len(imgs), 3, H, W, dtype=torch.uint8)
    for i, im in enumerate(imgs):
        batch_u8[i, :, : im.shape[1], : im.shape[2]] = im
    eng.forward_dense(batch_u8.cuda(), c["sizes"])
    torch.cuda.synchronize()
    batch, sizes = omodel.preprocess(imgs, spec.pixel_mean, spec.pixel_std)
    assert sizes == c["sizes"] and tuple(batch.shape[2:]) == (H, W)
    ref16 = omodel.forward_dense(sd, spec.resnet_depth, batch, "o16")
    ref32 = omodel.forward_dense(sd, spec.resnet_depth, batch, "fp32")
    yield request.param, eng, spec, sizes, ref16, ref32
    eng.close()


def _post(spec, logits, reg, ctr, sizes, osz=None):
    return opost.postprocess(logits, reg, ctr, spec.fpn_strides, sizes, osz, score_thresh=spec.score_thresh,
                             pre_nms_topk=spec.pre_nms_topk, nms_thresh=spec.nms_thresh,
                             post_nms_topk=spec.post_nms_topk, sort_corners=spec.sort_corners,
                             thresh_with_ctr=spec.thresh_with_ctr, vehicle_merge=spec.vehicle_merge)


def _gpu_heads(eng):
    heads = [eng.head_outputs(l) for l in range(5)]
    logits = [h["logits"].cpu().numpy() for h in heads]
    ctr = [h["ctr_delta"][:, :1].cpu().numpy() for h in heads]
    reg = [(np.tile(h["center"].cpu().numpy(), (1, 4, 1, 1)) + h["ctr_delta"][:, 1:9].cpu().numpy()).astype(np.float32)
           for h in heads]
    return logits, reg, ctr


def test_layerwise_vs_quantisation_matched_oracle(case):
    name, eng, spec, sizes, ref16, ref32 = case
    worst, worst_name = 0.0, ""
    for key, r in ref16["named"].items():
        if key == "stem":
            continue  # the stem's max-pool runs in its epilogue: the conv output exists only as "pool"
        a = eng.activation(key).cpu()
        assert a.shape == r.shape, key
        rel = ((a - r).norm() / (r.norm() + 1e-12)).item()
        if rel > worst:
            worst, worst_name = rel, key
        assert rel <= 8e-3, f"{name} {key}: rel L2 {rel}"
    assert worst > 0
    _record(name, "layerwise_worst_rel_l2_vs_o16", {"value": worst, "activation": worst_name})


def test_head_outputs_vs_fp32_reference_arithmetic(case):
    name, eng, spec, sizes, ref16, ref32 = case
    logits, reg, ctr = _gpu_heads(eng)
    d = {"logits": 0.0, "ctr": 0.0, "reg": 0.0}
    for l in range(5):
        d["logits"] = max(d["logits"], float(np.abs(logits[l] - ref32["logits"][l].numpy()).max()))
        d["ctr"] = max(d["ctr"], float(np.abs(ctr[l] - ref32["ctr"][l].numpy()).max()))
        d["reg"] = max(d["reg"], float(np.abs(reg[l] - ref32["reg"][l].numpy()).max()))
    _record(name, "head_max_abs_diff_vs_fp32", d)
    assert d["logits"] <= 3e-2 and d["ctr"] <= 3e-2 and d["reg"] <= 5e-2, d


def test_postprocess_bit_exact_on_gpu_heads(case):
    """The whole post-processing at the BASELINE shape (21 824 locations per image, the per-level top-k cap and the
    post-NMS cut of the YAML: 2000 / 1000) on the GPU's own head outputs: identical to the oracle, bit for bit."""
    name, eng, spec, sizes, ref16, ref32 = case
    logits, reg, ctr = _gpu_heads(eng)
    osz = [(h + 37, w - 11) for h, w in sizes]
    want = _post(spec, logits, reg, ctr, sizes, osz)
    dets, counts = eng.postprocess(sizes, osz, True)
    dets, counts = dets.cpu().numpy(), counts.cpu().numpy()
    n_in = [c["nms_in"] for c in eng.post_counts()]
    assert max(n_in) > 500, f"{name}: the case should give the NMS real work, got {n_in}"
    for i, w in enumerate(want):
        n = len(w["scores"])
        assert counts[i] == n, (name, i, counts[i], n)
        g = dets[i, :n]
        assert np.array_equal(g[:, 18].view(np.uint32).astype(np.int64), w["canon"])
        assert np.array_equal(g[:, 14].astype(np.int64), w["pred_classes"])
        assert np.array_equal(g[:, 0:8], w["pred_corners"])
        assert np.array_equal(g[:, 8:12], w["pred_boxes"])
        assert np.array_equal(g[:, 12], w["scores"])
        assert np.array_equal(g[:, 16:18], w["locations"])
    _record(name, "postprocess", {"nms_boxes_in": n_in, "detections": [int(c) for c in counts],
                                  "bit_exact_vs_oracle_on_gpu_heads": True})


def _agreement(a, b, strides):
    """How much of result `b` (list of per-image oracle dicts: the yardstick) is reproduced by `a`: matched by canonical
    candidate index; max |d score| and max |d coord| (in units of the level's stride) over the matched rows."""
    tot = hit = 0
    dscore = dcoord = 0.0
    for x, y in zip(a, b):
        pos = {int(c): k for k, c in enumerate(x["canon"])}
        idx = [(pos[int(c)], k) for k, c in enumerate(y["canon"]) if int(c) in pos]
        tot += len(y["canon"])
        hit += len(idx)
        if idx:
            ia, ib = np.array([p for p, _ in idx]), np.array([q for _, q in idx])
            dscore = max(dscore, float(np.abs(x["scores"][ia] - y["scores"][ib]).max()))
            # the same polygon may come out in another vertex order (sort_quadrilateral's strict comparisons): compare
            # as point sets through the hbox, which is order-free
            st = np.asarray(strides, np.float32)[y["fpn_levels"][ib]]
            dcoord = max(dcoord, float((np.abs(x["pred_boxes"][ia] - y["pred_boxes"][ib]).max(1) / st).max()))
    return {"reference_detections": tot, "reproduced": hit, "share": hit / max(tot, 1), "max_abs_dscore": dscore,
            "max_abs_dbox_in_strides": dcoord}


def test_end_to_end_agreement_is_measured_and_bounded(case):
    """Thresholds, top-k membership, IoU > 0.1 and the post-NMS cut are discontinuous: with ~10^3 candidates per image
    a few always sit inside the fp16 drift of a boundary, so identity with the fp32 reference is not attainable at
    this density (tests/test_identity_gpu.py gates identity on inputs chosen to have margins). Here the agreement is
    MEASURED -- and the kernel path must agree with the quantisation-matched oracle at least as well as that oracle
    agrees with fp32, i.e. the disagreement is the fp16 storage format's, not the kernels'."""
    name, eng, spec, sizes, ref16, ref32 = case
    logits, reg, ctr = _gpu_heads(eng)
    r_gpu = _post(spec, logits, reg, ctr, sizes)
    r16 = _post(spec, *[[t.numpy() for t in ref16[k]] for k in ("logits", "reg", "ctr")], sizes)
    r32 = _post(spec, *[[t.numpy() for t in ref32[k]] for k in ("logits", "reg", "ctr")], sizes)
    rep = {"kernel_vs_o16": _agreement(r_gpu, r16, spec.fpn_strides),
           "kernel_vs_fp32": _agreement(r_gpu, r32, spec.fpn_strides),
           "o16_vs_fp32": _agreement(r16, r32, spec.fpn_strides)}
    _record(name, "end_to_end", rep)
    assert rep["kernel_vs_fp32"]["reference_detections"] > 100
    assert rep["kernel_vs_o16"]["share"] >= 0.93, rep
    assert rep["kernel_vs_fp32"]["share"] >= 0.90, rep
    assert rep["kernel_vs_fp32"]["share"] >= rep["o16_vs_fp32"]["share"] - 0.03, rep
    assert rep["kernel_vs_fp32"]["max_abs_dscore"] <= 2e-3, rep
    assert rep["kernel_vs_fp32"]["max_abs_dbox_in_strides"] <= 0.12, rep
