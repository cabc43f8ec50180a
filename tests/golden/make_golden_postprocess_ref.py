"""Golden vectors for the inference post-processing chain, produced by the REFERENCE's own code. Run HERE (the
container that has /root/reference), never on the GPU box:

    python tests/golden/make_golden_postprocess_ref.py

What is executed is the reference's source, imported from where it lies (nothing is copied into this repository; only
inputs and outputs are stored in tests/golden/postprocess_ref_*.npz):

  DAFNeOutputs.predict_proposals / forward_for_single_feature_map / select_over_all_levels
                                          /root/reference/dafne/modeling/dafne/dafne_outputs.py:733-925
  ml_nms / batched_nms_poly               /root/reference/dafne/modeling/nms/nms.py:10-92
  sort_quadrilateral                      /root/reference/dafne/utils/sort_corners.py:26-92
  OneStageDetector.forward / ._postprocess /root/reference/dafne/modeling/one_stage_detector.py:45-98

The packages those files import but that cannot be installed here (no network) are replaced by stand-ins registered in
sys.modules BEFORE the import, so the reference files run unmodified:

  detectron2.structures.Instances / Boxes   -> dafne_b200.structures (same fields / indexing / cat semantics)
  detectron2.modeling.ProposalNetwork       -> `ProposalNetworkStandIn` below: the inference tail of detectron2 v0.5
                                               ProposalNetwork.forward [EXT, restated]: detector_postprocess(results,
                                               input.get("height", image_size[0]), input.get("width", image_size[1]))
  detectron2.modeling.postprocessing.detector_postprocess -> `d2_detector_postprocess` below [EXT, restated]: scale
                                               pred_boxes, clip to the output size, drop rows whose box is empty
  poly_nms.poly_gpu_nms                     -> argsort(scores)[::-1] + the greedy sweep over IoU > thresh with the float
                                               instantiation of the reference's polyiou.cpp arithmetic
                                               (oracle/polyiou_oracle.c; its double twin is pinned against the
                                               reference's polyiou.cpp itself, tests/test_oracle_cpu.py)
  fvcore, detectron2.utils.*, the loss modules (training only)  -> empty shells
  the `dafne`, `dafne.modeling`, ... package __init__ files (they import every backbone)  -> shells with the real __path__

The head outputs are synthetic (seeded); every case is checked for exact score ties before it is stored, because the
reference leaves tie order to `topk(sorted=False)` / numpy `argsort`.
"""
import importlib
import os
import sys
import types

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from dafne_b200.structures import Boxes, Instances  # noqa: E402
from oracle import postprocess as opost  # noqa: E402


# ------------------------------------------------------------------------------------------ detectron2 stand-ins
def d2_detector_postprocess(results, output_height, output_width, mask_threshold=0.5):
    """detectron2 v0.5 modeling/postprocessing.py::detector_postprocess for box-only results [EXT, restated]."""
    new_size = (output_height, output_width)
    scale_x, scale_y = (output_width / results.image_size[1], output_height / results.image_size[0])
    results = Instances(new_size, **results.get_fields())
    output_boxes = results.pred_boxes
    output_boxes.scale(scale_x, scale_y)
    output_boxes.clip(results.image_size)
    results = results[output_boxes.nonempty()]
    return results


class ProposalNetworkStandIn(nn.Module):
    """The inference tail of detectron2 v0.5 ProposalNetwork.forward [EXT, restated]. The dense part (normalise, pad,
    backbone, head) is not under test here: `proposal_fn(batched_inputs)` hands over what `proposal_generator` returned
    -- the reference's own predict_proposals output -- together with ImageList.image_sizes."""

    def __init__(self, cfg=None):
        super().__init__()

    def forward(self, batched_inputs):
        proposals, image_sizes = self.proposal_fn(batched_inputs)
        processed_results = []
        for results_per_image, input_per_image, image_size in zip(proposals, batched_inputs, image_sizes):
            height = input_per_image.get("height", image_size[0])
            width = input_per_image.get("width", image_size[1])
            r = d2_detector_postprocess(results_per_image, height, width)
            processed_results.append({"proposals": r})
        return processed_results


class _Registry:
    def register(self, obj=None):
        return (lambda o: o) if obj is None else obj


def poly_gpu_nms(dets, thresh, device_id=0):
    """External poly_nms.poly_gpu_nms [EXT, restated]: order = scores.argsort()[::-1]; keep i unless an earlier kept j
    has IoU(j, i) > thresh; return order[keep]."""
    dets = np.ascontiguousarray(dets, np.float32)
    order = dets[:, 8].argsort()[::-1]
    keep = opost.greedy_nms(dets[order, :8], float(thresh))
    return [int(v) for v in order[keep]]


def install_stand_ins():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    for pkg in ("dafne", "dafne.modeling", "dafne.modeling.dafne", "dafne.modeling.nms", "dafne.utils"):
        mod(pkg).__path__ = [os.path.join(REF, *pkg.split("."))]
    mod("dafne.modeling.losses").__path__ = []
    mod("dafne.modeling.losses.smooth_l1", ModulatedEightPointLoss=object, SmoothL1Loss=object)
    mod("fvcore").__path__ = []
    mod("fvcore.nn", sigmoid_focal_loss_jit=None)
    mod("detectron2").__path__ = []
    mod("detectron2.layers", cat=torch.cat)
    mod("detectron2.structures", Instances=Instances, Boxes=Boxes, ImageList=None).__path__ = []
    mod("detectron2.structures.boxes", Boxes=Boxes)
    mod("detectron2.utils").__path__ = []
    mod("detectron2.utils.comm", get_world_size=lambda: 1, get_local_rank=lambda: 0)
    mod("detectron2.utils.events", get_event_storage=None)
    mod("detectron2.utils.logger", log_first_n=None)
    mod("detectron2.modeling", ProposalNetwork=ProposalNetworkStandIn, GeneralizedRCNN=nn.Module).__path__ = []
    mod("detectron2.modeling.meta_arch").__path__ = []
    mod("detectron2.modeling.meta_arch.build", META_ARCH_REGISTRY=_Registry())
    mod("detectron2.modeling.postprocessing", detector_postprocess=d2_detector_postprocess)
    mod("poly_nms", poly_gpu_nms=poly_gpu_nms)
    mod("poly_overlaps", poly_overlaps=None)


def reference_objects(num_classes, sort_corners, thresh_with_ctr, pre_topk, post_topk, strides):
    """(DAFNeOutputs instance, OneStageDetector instance) of the reference, configured like the YAMLs configure them."""
    install_stand_ins()
    outputs_mod = importlib.import_module("dafne.modeling.dafne.dafne_outputs")
    osd_mod = importlib.import_module("dafne.modeling.one_stage_detector")
    assert outputs_mod.__file__.startswith(REF) and osd_mod.__file__.startswith(REF)
    assert sys.modules["dafne.modeling.nms.nms"].__file__.startswith(REF)
    assert sys.modules["dafne.utils.sort_corners"].__file__.startswith(REF)
    ns = types.SimpleNamespace
    o = outputs_mod.DAFNeOutputs.__new__(outputs_mod.DAFNeOutputs)
    nn.Module.__init__(o)
    o.cfg = ns(MODEL=ns(DAFNE=ns(ENABLE_FPN_STRIDE_NORM=True)))
    o.pre_nms_thresh_test, o.pre_nms_topk_test, o.post_nms_topk_test = 0.05, pre_topk, post_topk
    o.nms_thresh = 0.1
    o.thresh_with_ctr = thresh_with_ctr
    o.has_centerness = True  # CENTERNESS: oriented / plain in every shipped config
    o.sort_corners = sort_corners
    o.num_classes = num_classes
    o.strides = list(strides)
    o.eval()
    det = osd_mod.OneStageDetector.__new__(osd_mod.OneStageDetector)
    nn.Module.__init__(det)
    det.eval()
    return o, det


# ------------------------------------------------------------------------------------------ synthetic head outputs
def synth_heads(seed, N, C, Hs, Ws, cls_bias, base_quad=(-3, -1, 3, -1, 3, 1, -3, 1), reg_noise=1.0):
    rng = np.random.default_rng(seed)
    logits, reg, ctr = [], [], []
    for h, w in zip(Hs, Ws):
        logits.append(rng.normal(cls_bias, 1.3, (N, C, h, w)).astype(np.float32))
        base = np.asarray(base_quad, np.float32).reshape(1, 8, 1, 1)
        reg.append((base + rng.normal(0, reg_noise, (N, 8, h, w))).astype(np.float32))
        ctr.append(rng.normal(0, 1.0, (N, 1, h, w)).astype(np.float32))
    return logits, reg, ctr


def compute_locations(h, w, stride):
    """dafne/modeling/dafne/dafne.py:37-44, called the way DAFNe.compute_locations (:158-164) calls it."""
    shifts_x = torch.arange(0, w * stride, step=stride, dtype=torch.float32)
    shifts_y = torch.arange(0, h * stride, step=stride, dtype=torch.float32)
    shift_y, shift_x = torch.meshgrid(shifts_y, shifts_x, indexing="ij")
    return torch.stack((shift_x.reshape(-1), shift_y.reshape(-1)), dim=1) + stride // 2


CASES = {
    # tag: C, SORT_CORNERS, THRESH_WITH_CTR, pre_nms_topk, post_nms_topk, seed, cls_bias, sizes / outputs per image
    # dota-1.0 1024.yaml flavour; the top-k cap binds at the first two levels, the post-NMS cut binds
    "c15_sort": dict(C=15, sort=True, twc=False, pre=400, post=150, seed=11, bias=-3.2),
    # dota-1.0_r101_ms flavour: threshold on sqrt(cls * ctr)
    "c15_ctr": dict(C=15, sort=True, twc=True, pre=400, post=150, seed=12, bias=-2.2),
    # dota-1.5 flavour (16 classes, unsorted corners)
    "c16_nosort": dict(C=16, sort=False, twc=False, pre=400, post=150, seed=13, bias=-3.2),
    # hrsc_r50_ms flavour: one class, 8:1 quads
    "c1_sort": dict(C=1, sort=True, twc=False, pre=400, post=150, seed=14, bias=-1.2,
                    quad=(-4, -0.5, 4, -0.5, 4, 0.5, -4, 0.5)),
    # the YAML values themselves (2000 / 1000): neither cut binds
    "c15_default_topk": dict(C=15, sort=True, twc=False, pre=2000, post=1000, seed=15, bias=-3.6),
}
HS, WS, STRIDES = [24, 12, 6, 3, 2], [32, 16, 8, 4, 2], [8, 16, 32, 64, 128]


def run_reference(cfg, do_postprocess, tag="live"):
    """Synthetic head outputs of one case through the reference's code: (fixture dict without outputs, per-image results)."""
    o, det = reference_objects(cfg["C"], cfg["sort"], cfg["twc"], cfg["pre"], cfg["post"], STRIDES)
    N = 3
    logits, reg, ctr = synth_heads(cfg["seed"], N, cfg["C"], HS, WS, cfg["bias"],
                                   cfg.get("quad", (-3, -1, 3, -1, 3, 1, -3, 1)))
    H, W = HS[0] * 8, WS[0] * 8
    sizes = [(H, W), (H - 20, W - 12), (H - 64, W - 96)]  # ImageList.image_sizes: un-padded sizes
    outs = [(2 * H, 2 * W), None, (H - 64 + 37, W - 96 - 11)]  # "height"/"width" of the input dicts (None: absent)
    locations = [compute_locations(h, w, s) for h, w, s in zip(HS, WS, STRIDES)]

    def proposal_fn(batched_inputs):
        props = o.predict_proposals([torch.from_numpy(a) for a in logits], [torch.from_numpy(a) for a in reg],
                                    [torch.from_numpy(a) for a in ctr], locations, sizes, top_feats=[])
        return props, sizes

    det.proposal_fn = proposal_fn
    inputs = []
    for (h, w), out in zip(sizes, outs):
        d = {"image": torch.zeros(3, h, w)}
        if out is not None:
            d["height"], d["width"] = out
        inputs.append(d)
    if do_postprocess:
        # the reference's _postprocess reads inp["height"] unconditionally (one_stage_detector.py:84): every dict has it
        for d, (h, w) in zip(inputs, sizes):
            d.setdefault("height", h)
            d.setdefault("width", w)
    with torch.no_grad():
        res = det.forward(inputs, do_postprocess=do_postprocess)
    blob = {}
    for l in range(5):
        blob[f"logits{l}"], blob[f"reg{l}"], blob[f"ctr{l}"] = logits[l], reg[l], ctr[l]
    osz = [(d.get("height", s[0]), d.get("width", s[1])) for d, s in zip(inputs, sizes)]
    blob["sizes"] = np.array(sizes)
    blob["osz"] = np.array(osz)
    blob["meta"] = np.array([cfg["C"], int(cfg["sort"]), int(cfg["twc"]), cfg["pre"], cfg["post"], int(do_postprocess)])
    results = []
    for i, r in enumerate(res):
        inst = r["instances"]
        assert inst.image_size == tuple(osz[i])
        sc = inst.scores.numpy()
        assert len(np.unique(sc)) == len(sc), f"{tag}: exact score ties in image {i}; pick another seed"
        results.append(dict(pred_corners=inst.pred_corners.numpy(), pred_boxes=inst.pred_boxes.tensor.numpy(), scores=sc,
                            centerness=inst.centerness.numpy(), pred_classes=inst.pred_classes.numpy(),
                            fpn_levels=inst.fpn_levels.numpy(), locations=inst.locations.numpy()))
    return blob, results


def run_case(tag, cfg, do_postprocess):
    blob, results = run_reference(cfg, do_postprocess, tag)
    for i, r in enumerate(results):
        for k, v in r.items():
            blob[f"out{i}_{k}"] = v
    name = f"postprocess_ref_{tag}" + ("" if do_postprocess else "_nopost") + ".npz"
    np.savez_compressed(os.path.join(HERE, name), **blob)
    print(name, [len(r["scores"]) for r in results])


def main():
    for tag, cfg in CASES.items():
        run_case(tag, cfg, True)
    run_case("c15_sort", CASES["c15_sort"], False)


if __name__ == "__main__":
    main()
