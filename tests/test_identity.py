"""The end-to-end IDENTITY gate: on fixed synthetic inputs the CUDA path returns the same detections as the reference's
fp32 arithmetic (oracle/model.py "fp32" + oracle/postprocess.py) -- the same candidate indices and class ids after NMS,
scores within 1e-3, coordinates within the fp16 storage format's reach.

Why the inputs are special: the reference computes the dense forward in fp32,
this repository stores fp16 activations between ~100 layers; the head outputs drift by |d logit| <= 1e-2 (RMS 1.6e-3),
and every discontinuous decision of the post-processing -- score > 0.05, membership of the per-level top-k, IoU > 0.1,
the post-NMS cut -- flips for a candidate that sits inside that drift of its boundary. At the density of the bench
workloads (10^3 candidates per image) some always do (measured agreement 97-99 %, tests/test_baseline_shapes_gpu.py).
The inputs here are CHOSEN so that none does, by scripts/find_identity_input.py:
  * image seed, size 256 x 320, synthetic weights seed 0;
  * the 15 per-class biases of cls_logits sit in the middle of the widest gap of that class's threshold-crossing values
    (half-gap >= 0.009 logit units = 5 x the RMS drift): that places the score threshold;
  * the seed was accepted after the fp32 oracle, the quantisation-matched oracle (o16) and 12 runs of the fp32 heads
    perturbed by noise of 3 x the RMS drift all returned the same detections: NMS and ordering have margins too.
"Identical" = the same set of (candidate index, class id) per image; two detections whose fp32 scores differ by less
than the drift may swap places in the score-sorted output, so the ORDER is compared wherever the fp32 gap is > 2e-3.
"""
import numpy as np
import pytest
import torch

from oracle import model as omodel
from oracle import postprocess as opost

CASES = {
    # dota-1.0 1024.yaml flavour: R50, C = 15, SORT_CORNERS, threshold on the class score
    "r50": dict(depth=50, twc=False, seed=1000, hw=(256, 320), biases=[
        -3.587542772293091, -4.205982685089111, -4.460017681121826, -4.302076816558838, -4.689307689666748,
        -3.304135322570801, -3.9970812797546387, -4.380699157714844, -3.3084826469421387, -3.0946309566497803,
        -3.709336042404175, -4.282752513885498, -4.058029651641846, -4.152675628662109, -3.7173454761505127]),
    # dota-1.0_r101_ms flavour (the headline config): R101, THRESH_WITH_CTR
    "r101_ctr": dict(depth=101, twc=True, seed=2000, hw=(256, 320), biases=[
        -5.758866786956787, -6.257762432098389, -6.301433086395264, -7.087118625640869, -5.907781600952148,
        -7.163401126861572, -6.136742115020752, -7.307214260101318, -6.3969807624816895, -6.924020290374756,
        -6.240398406982422, -6.208432674407959, -6.6732587814331055, -6.051267147064209, -5.498477458953857]),
}
_CACHE = {}


def _setup(name):
    """(spec, state dict, image, sizes, fp32 oracle detections) of a case; the oracle runs once per process."""
    if name in _CACHE:
        return _CACHE[name]
    from dafne_b200.spec import ModelSpec
    from dafne_b200.weights import synthetic_state_dict

    c = CASES[name]
    spec = ModelSpec(resnet_depth=c["depth"], num_classes=15, thresh_with_ctr=c["twc"])
    sd = synthetic_state_dict(spec, 0)
    sd["proposal_generator.dafne_head.cls_logits.bias"] = torch.tensor(c["biases"], dtype=torch.float32)
    g = torch.Generator().manual_seed(c["seed"])
    img = torch.randint(0, 256, (3, *c["hw"]), dtype=torch.uint8, generator=g)
    batch, sizes = omodel.preprocess([img], spec.pixel_mean, spec.pixel_std)
    _CACHE[name] = (spec, sd, img, sizes, batch)
    return _CACHE[name]


def _post(spec, out, sizes):
    return opost.postprocess([t.numpy() for t in out["logits"]], [t.numpy() for t in out["reg"]],
                             [t.numpy() for t in out["ctr"]], spec.fpn_strides, sizes, None,
                             score_thresh=spec.score_thresh, pre_nms_topk=spec.pre_nms_topk, nms_thresh=spec.nms_thresh,
                             post_nms_topk=spec.post_nms_topk, sort_corners=spec.sort_corners,
                             thresh_with_ctr=spec.thresh_with_ctr, vehicle_merge=spec.vehicle_merge)[0]


def _assert_identical(got, want, strides, what):
    """got / want: dicts with canon, pred_classes, scores, pred_corners, fpn_levels (one image)."""
    assert len(got["canon"]) == len(want["canon"]), (what, len(got["canon"]), len(want["canon"]))
    og, ow = np.argsort(got["canon"]), np.argsort(want["canon"])
    assert np.array_equal(got["canon"][og], want["canon"][ow]), f"{what}: different candidate indices survive"
    assert np.array_equal(got["pred_classes"][og], want["pred_classes"][ow]), f"{what}: class ids differ"
    dscore = float(np.abs(got["scores"][og] - want["scores"][ow]).max())
    assert dscore <= 1e-3, (what, dscore)
    # coordinates: the head regresses in stride units and fp16 activations carry 11 bits, so the reach is a fraction of
    # a stride (measured <= 0.02), i.e. <= 1e-3 of the 1024-pixel extent up to p5 and 2.5e-3 at p7 (stride 128)
    st = np.asarray(strides, np.float32)[want["fpn_levels"][ow]]
    gs, ws = got["pred_corners"][og].reshape(-1, 4, 2), want["pred_corners"][ow].reshape(-1, 4, 2)
    # sort_quadrilateral starts at the leftmost vertex by a strict comparison: compare as point sets
    d = np.abs(np.sort(gs, axis=1) - np.sort(ws, axis=1)).reshape(len(st), -1).max(1) / st
    assert float(d.max()) <= 0.03, (what, float(d.max()))
    # order: wherever the reference's neighbours in the score-sorted list are more than the drift apart
    pos = {int(c): k for k, c in enumerate(got["canon"])}
    for k in range(len(want["canon"]) - 1):
        if want["scores"][k] - want["scores"][k + 1] > 2e-3:
            assert pos[int(want["canon"][k])] < pos[int(want["canon"][k + 1])], (what, "order", k)
    return dscore, float(d.max())


@pytest.mark.parametrize("name", list(CASES))
def test_documented_inputs_have_margins_on_the_cpu(name):
    """The inputs are what the header says they are: the quantisation-matched oracle (fp16 storage between layers, the
    model of what the kernels compute) returns the same detections as the fp32 oracle."""
    spec, sd, img, sizes, batch = _setup(name)
    r32 = _post(spec, omodel.forward_dense(sd, spec.resnet_depth, batch, "fp32"), sizes)
    r16 = _post(spec, omodel.forward_dense(sd, spec.resnet_depth, batch, "o16"), sizes)
    assert len(r32["scores"]) >= 30
    _assert_identical(r16, r32, spec.fpn_strides, name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_cuda_path_returns_the_reference_detections(name):
    from dafne_b200.engine import DafneEngine

    spec, sd, img, sizes, batch = _setup(name)
    want = _post(spec, omodel.forward_dense(sd, spec.resnet_depth, batch, "fp32"), sizes)
    eng = DafneEngine(spec, torch.device("cuda:0"))
    eng.load_state_dict(sd)
    dets, counts = eng.detect(img[None].cuda(), sizes)
    n = int(counts[0])
    g = dets[0, :n].cpu().numpy()
    got = dict(canon=g[:, 18].view(np.uint32).astype(np.int64), pred_classes=g[:, 14].astype(np.int64), scores=g[:, 12],
               pred_corners=g[:, 0:8], fpn_levels=g[:, 15].astype(np.int64))
    dscore, dcoord = _assert_identical(got, want, spec.fpn_strides, name)
    print(f"{name}: {n} detections identical to the fp32 reference arithmetic; max |d score| {dscore:.2e}, "
          f"max |d corner| {dcoord:.3f} strides")
    eng.close()
