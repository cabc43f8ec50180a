"""The reference's Python surface for the inference path, backed by the sm_100a kernels.

Mirrors (names, argument meaning, return types, error behaviour):
  OneStageDetector(cfg).forward / .inference / ._postprocess      dafne/modeling/one_stage_detector.py:34-107
  model.proposal_generator.dafne_outputs.select_over_all_levels     dafne/modeling/dafne/dafne_outputs.py:907-925 (used by TTA, tta.py:265-267)
  ml_nms, batched_nms_poly                                          dafne/modeling/nms/nms.py:10-92
  poly_gpu_nms(dets, thresh, device_id)                             external poly_nms module (nms.py:6,91)
  sort_quadrilateral                                                dafne/utils/sort_corners.py:26-92
  META_ARCH_REGISTRY / PROPOSAL_GENERATOR_REGISTRY / BACKBONE_REGISTRY + build_model(cfg)   detectron2 registries

Everything numeric goes through libdafne_b200.so; torch is device memory, streams and the nn.Module shell that
gives the model a detectron2-compatible state dict.
"""
from __future__ import annotations

from collections import OrderedDict
import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
from torch import nn

from . import _capi
from .engine import DET, DafneEngine
from .spec import ModelSpec
from .structures import Boxes, Instances
from .weights import state_dict_shapes, synthetic_state_dict


class Registry(dict):
    def __init__(self, name: str):
        super().__init__()
        self._name = name

    def register(self, obj=None):
        def deco(o):
            self[o.__name__] = o
            return o

        return deco if obj is None else deco(obj)

    def get(self, name):  # type: ignore[override]
        if name not in self:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
        return self[name]


META_ARCH_REGISTRY = Registry("META_ARCH")
PROPOSAL_GENERATOR_REGISTRY = Registry("PROPOSAL_GENERATOR")
BACKBONE_REGISTRY = Registry("BACKBONE")


class _ParamTree(nn.Module):
    """Nested module whose state-dict keys are exactly the dotted names given (e.g. backbone.bottom_up.res2.0...)."""

    def add(self, dotted: str, tensor: torch.Tensor) -> None:
        head, _, rest = dotted.partition(".")
        if not rest:
            self.register_buffer(head, tensor)
            return
        if head not in self._modules:
            self.add_module(head, _ParamTree())
        self._modules[head].add(rest, tensor)


# ------------------------------------------------------------------------------------------------ functional ops
def _scratch(nbytes: int, device) -> torch.Tensor:
    return torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)


def _aligned(t: torch.Tensor) -> int:
    return (t.data_ptr() + 1023) // 1024 * 1024


def sort_quadrilateral(bboxes: torch.Tensor) -> torch.Tensor:
    """Canonical corner order of quadrilaterals [n, 8] (sort_corners.py:26-92) on the GPU."""
    assert bboxes.dim() == 2
    if bboxes.shape[0] == 0:
        return bboxes
    if not bboxes.is_cuda:
        raise _capi.DafneError("sort_quadrilateral: CUDA tensor required (no CPU fallback in this package)")
    x = bboxes.contiguous().float()
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _capi.check(_capi.lib().dafne_sort_quadrilateral(x.data_ptr(), out.data_ptr(), x.shape[0], _capi.stream_ptr()),
                    "dafne_sort_quadrilateral")
    return out


def poly_iou(p: torch.Tensor, q: torch.Tensor) -> torch.Tensor:
    """Row-wise polygon IoU of quadrilaterals [n, 8] in the faithful fp32 arithmetic (polyiou.cpp:108-133)."""
    p = p.contiguous().float()
    q = q.contiguous().float()
    out = torch.empty(p.shape[0], dtype=torch.float32, device=p.device)
    with torch.cuda.device(p.device):
        _capi.check(_capi.lib().dafne_poly_iou(p.data_ptr(), q.data_ptr(), out.data_ptr(), p.shape[0],
                                               _capi.stream_ptr()), "dafne_poly_iou")
    return out


def resize_bilinear_u8(image: torch.Tensor, new_h: int, new_w: int) -> torch.Tensor:
    """[C, H, W] (or [N, C, H, W]) uint8 CUDA tensor -> same layout at new_h x new_w, bit-identical to
    PIL.Image.resize((new_w, new_h), Image.BILINEAR) -- what detectron2's ResizeTransform runs on the host
    (tools/plain_train_net.py:293-298, dafne/modeling/tta.py:76-93)."""
    if not image.is_cuda or image.dtype != torch.uint8:
        raise _capi.DafneError("resize_bilinear_u8: uint8 CUDA tensor required (no CPU fallback in this package)")
    x = image.contiguous()
    h, w = int(x.shape[-2]), int(x.shape[-1])
    planes = x.numel() // (h * w)
    out = torch.empty(tuple(x.shape[:-2]) + (int(new_h), int(new_w)), dtype=torch.uint8, device=x.device)
    tmp = torch.empty(planes * h * int(new_w), dtype=torch.uint8, device=x.device) if (new_h != h and new_w != w) else None
    with torch.cuda.device(x.device):
        _capi.check(_capi.lib().dafne_resize_bilinear_u8(x.data_ptr(), planes, h, w, out.data_ptr(), int(new_h),
                                                         int(new_w), tmp.data_ptr() if tmp is not None else None,
                                                         _capi.stream_ptr()), "dafne_resize_bilinear_u8")
    return out


def batched_nms_poly(boxes: torch.Tensor, scores: torch.Tensor, idxs: torch.Tensor, iou_threshold: float,
                     vehicle_merge: bool = True) -> torch.Tensor:
    """Class-aware polygon NMS (nms.py:37-92): int64 indices of kept boxes, in decreasing score order."""
    assert boxes.shape[-1] == 8
    if boxes.numel() == 0:
        return torch.empty((0,), dtype=torch.int64, device=boxes.device)
    keep, nkeep, _ = _batched_nms_poly_launch(boxes, scores, idxs, iou_threshold, vehicle_merge)
    return keep[: int(nkeep.item())].to(torch.int64)


def _batched_nms_poly_launch(boxes: torch.Tensor, scores: torch.Tensor, idxs: torch.Tensor, iou_threshold: float,
                             vehicle_merge: bool = True):
    """Enqueue the NMS of `batched_nms_poly` on the current stream WITHOUT reading its result: (keep int32 [n], nkeep
    int32 [1], tensors the kernels still use). Callers that run several images overlap them on separate streams."""
    if not boxes.is_cuda:
        raise _capi.DafneError("batched_nms_poly: CUDA tensors required (no CPU fallback in this package)")
    n = boxes.shape[0]
    lib = _capi.lib()
    b = boxes.contiguous().float()
    s = scores.contiguous().float()
    c = idxs.to(torch.int32).contiguous()
    need = C.c_size_t()
    _capi.check(lib.dafne_poly_nms_scratch_bytes(n, C.byref(need)), "dafne_poly_nms_scratch_bytes")
    scratch = _scratch(need.value, boxes.device)
    keep = torch.empty(n, dtype=torch.int32, device=boxes.device)
    nkeep = torch.zeros(1, dtype=torch.int32, device=boxes.device)
    with torch.cuda.device(boxes.device):
        _capi.check(lib.dafne_poly_nms(b.data_ptr(), s.data_ptr(), c.data_ptr(), n, float(iou_threshold),
                                       int(vehicle_merge), keep.data_ptr(), nkeep.data_ptr(), _aligned(scratch),
                                       need.value, _capi.stream_ptr()), "dafne_poly_nms")
    return keep, nkeep, (b, s, c, scratch)


def ml_nms(boxlist: Instances, nms_thresh: float, max_proposals: int = -1) -> Instances:
    """nms.py:10-33."""
    if nms_thresh <= 0:
        return boxlist
    if boxlist.scores.shape[0] == 0:
        return boxlist
    keep = batched_nms_poly(boxlist.pred_corners, boxlist.scores, boxlist.pred_classes, nms_thresh)
    if max_proposals > 0:
        keep = keep[:max_proposals]
    return boxlist[keep]


def poly_gpu_nms(dets: np.ndarray, thresh: float, device_id: int = 0) -> List[int]:
    """Drop-in for the external `poly_nms.poly_gpu_nms` the reference calls at nms.py:91: host numpy [n, 9] in
    (8 coordinates with class offsets already applied + score), list of kept indices out, descending score."""
    dets = np.ascontiguousarray(dets, dtype=np.float32)
    n = dets.shape[0]
    if n == 0:
        return []
    keep = (C.c_int * n)()
    num = C.c_int(0)
    _capi.check(_capi.lib().dafne_poly_nms_host(keep, C.byref(num), dets.ctypes.data_as(C.POINTER(C.c_float)), n,
                                                dets.shape[1], float(thresh), int(device_id)), "dafne_poly_nms_host")
    return list(keep[: num.value])


# ------------------------------------------------------------------------------------------------ module surface
class DAFNeOutputs:
    """The slice of dafne_outputs.DAFNeOutputs the inference callers touch."""

    def __init__(self, spec: ModelSpec):
        self.nms_thresh = spec.nms_thresh
        self.post_nms_topk = spec.post_nms_topk
        self.pre_nms_thresh = spec.score_thresh
        self.pre_nms_topk = spec.pre_nms_topk

    def select_over_all_levels(self, boxlists: List[Instances]) -> List[Instances]:
        """dafne_outputs.py:907-925: polygon NMS per image, then keep the post_nms_topk best (ties kept)."""
        results = []
        nms_results = self._ml_nms_overlapped(boxlists) if len(boxlists) > 1 else None
        for k, boxlist in enumerate(boxlists):
            result = nms_results[k] if nms_results is not None else ml_nms(boxlist, self.nms_thresh)
            n = len(result)
            if n > self.post_nms_topk > 0:
                scores = result.scores
                thr = scores[self.post_nms_topk - 1]  # rows are in descending score: the topk-th largest
                result = result[torch.nonzero(scores >= thr).squeeze(1)]
            results.append(result)
        return results


    _streams: List["torch.cuda.Stream"] = []

    def _ml_nms_overlapped(self, boxlists: List[Instances]) -> Optional[List[Instances]]:
        """`ml_nms` of several images at once: the NMS of one image is a chain of dependent launches that fills a
        fraction of the GPU (the greedy sweep inside a 512-box panel runs on one CTA), so the images run side by side on
        up to four streams and their results are read afterwards. None: nothing to overlap (the per-image path runs)."""
        if self.nms_thresh <= 0 or not all(b.scores.is_cuda and b.scores.shape[0] > 0 for b in boxlists):
            return None
        dev = boxlists[0].scores.device
        with torch.cuda.device(dev):
            cur = torch.cuda.current_stream()
            while len(DAFNeOutputs._streams) < min(4, len(boxlists)):
                DAFNeOutputs._streams.append(torch.cuda.Stream(device=dev))
            if any(st.device != dev for st in DAFNeOutputs._streams):
                DAFNeOutputs._streams[:] = [torch.cuda.Stream(device=dev) for _ in DAFNeOutputs._streams]
            pending = []
            for k, b in enumerate(boxlists):
                st = DAFNeOutputs._streams[k % len(DAFNeOutputs._streams)]
                st.wait_stream(cur)
                with torch.cuda.stream(st):
                    keep, nkeep, held = _batched_nms_poly_launch(b.pred_corners, b.scores, b.pred_classes, self.nms_thresh)
                for t in (keep, nkeep) + tuple(held):
                    t.record_stream(cur)
                pending.append((keep, nkeep, held))
            for st in DAFNeOutputs._streams:
                cur.wait_stream(st)
            counts = torch.cat([nk for _, nk, _ in pending]).tolist()  # one sync for all images
            return [b[keep[:n].to(torch.int64)] for b, (keep, _, _), n in zip(boxlists, pending, counts)]


@PROPOSAL_GENERATOR_REGISTRY.register()
class DAFNe:
    def __init__(self, spec: ModelSpec):
        self.in_features = ["p3", "p4", "p5", "p6", "p7"]
        self.fpn_strides = list(spec.fpn_strides)
        self.dafne_outputs = DAFNeOutputs(spec)


@BACKBONE_REGISTRY.register()
def build_dafne_resnet_fpn_backbone(cfg, input_shape=None):
    """The backbone is part of the fused launch plan; this entry exists so the registry name in the configs resolves."""
    return {"name": "resnet_fpn_p6p7", "depth": int(cfg.MODEL.RESNETS.DEPTH), "size_divisibility": 32}


@META_ARCH_REGISTRY.register()
class OneStageDetector(nn.Module):
    """`OneStageDetector(cfg)(batched_inputs) -> [{"instances": Instances}]`, inference only.

    batched_inputs: list of dicts with "image" (CHW uint8 or float tensor, the channel order MODEL.PIXEL_MEAN is
    given in), and optionally "height" / "width" = the resolution the outputs are rescaled to.
    """

    def __init__(self, cfg, init: str = "synthetic", seed: int = 0):
        super().__init__()
        self.cfg = cfg
        self.spec = ModelSpec.from_cfg(cfg)
        self.size_divisibility = self.spec.size_divisibility
        self.params = _ParamTree()
        if init == "synthetic":
            sd = synthetic_state_dict(self.spec, seed)
        else:
            sd = {k: torch.zeros(s) for k, s in state_dict_shapes(self.spec).items()}
        for k, v in sd.items():
            self.params.add(k, v)
        self.register_buffer("pixel_mean", torch.tensor(self.spec.pixel_mean).view(-1, 1, 1), persistent=False)
        self.register_buffer("pixel_std", torch.tensor(self.spec.pixel_std).view(-1, 1, 1), persistent=False)
        self.proposal_generator = DAFNe(self.spec)
        self.backbone = build_dafne_resnet_fpn_backbone(cfg)
        self.top_module = None  # MODEL.TOP_MODULE.NAME == "" in every shipped config (defaults.py:34)
        # One engine (context + bound workspace + launch plan) per batch shape, most recently used first: a caller
        # that cycles through shapes -- TTA runs 9 scales per image -- must not re-plan on every call. `_engine` is
        # the one used last.
        self._engine: Optional[DafneEngine] = None
        self._engines: "OrderedDict[tuple, DafneEngine]" = OrderedDict()
        self._engine_versions: Dict[int, int] = {}
        self._weights_version = 0
        self.max_cached_shapes = 12
        self.use_cuda_graphs = False  # opt-in (see _launch): pays off for callers that repeat shapes AND sizes
        self._weights_dirty = True
        self.eval()

    # -- state dict with detectron2 names -------------------------------------------------------------
    def _mark_dirty(self):
        self._weights_dirty = True

    def state_dict(self, *args, **kwargs):  # type: ignore[override]
        return self.params.state_dict(*args, **kwargs)

    def load_state_dict(self, state_dict, strict: bool = True):  # type: ignore[override]
        if "model" in state_dict and isinstance(state_dict["model"], dict):  # detectron2 checkpoint wrapper
            state_dict = state_dict["model"]
        state_dict = {k: torch.as_tensor(v) for k, v in state_dict.items()}
        res = self.params.load_state_dict(state_dict, strict=strict)
        self._weights_dirty = True
        return res

    def train(self, mode: bool = True):
        if mode:
            raise NotImplementedError("dafne_b200 implements the inference path only (SURVEY.md section 8)")
        return super().train(False)

    @property
    def device(self) -> torch.device:
        return self.pixel_mean.device

    def _get_engine(self, shape: Optional[tuple] = None) -> DafneEngine:
        """The engine bound (or to be bound) to batch shape (N, H, W); None = the one used last."""
        if not self.device.type == "cuda":
            raise _capi.DafneError("OneStageDetector must live on a CUDA device: call model.to('cuda') first")
        if self._weights_dirty:
            self._weights_version += 1
            self._weights_dirty = False
        if any(e.device != self.device for e in self._engines.values()):
            for e in self._engines.values():
                e.close()
            self._engines.clear()
            self._engine = None
        if shape is None:
            shape = next(reversed(self._engines)) if self._engines else ("default",)
        eng = self._engines.get(shape)
        if eng is None:
            if len(self._engines) >= self.max_cached_shapes:
                _, old = self._engines.popitem(last=False)
                self._engine_versions.pop(id(old), None)
                old.close()
            eng = DafneEngine(self.spec, self.device)
            self._engines[shape] = eng
        self._engines.move_to_end(shape)
        if self._engine_versions.get(id(eng)) != self._weights_version:
            eng.load_state_dict(self.params.state_dict())
            self._engine_versions[id(eng)] = self._weights_version
        self._engine = eng
        return eng

    # -- the reference's methods -----------------------------------------------------------------------
    def preprocess_image(self, batched_inputs: Sequence[dict]):
        """Batch the images (zero canvas, size divisibility 32); normalisation happens in the first kernel.
        Returns (N x 3 x H x W tensor on the device, [(h, w)])."""
        images = [x["image"] for x in batched_inputs]
        sizes = [(int(im.shape[1]), int(im.shape[2])) for im in images]
        d = self.size_divisibility
        H = (max(s[0] for s in sizes) + d - 1) // d * d
        W = (max(s[1] for s in sizes) + d - 1) // d * d
        dtype = torch.uint8 if all(im.dtype == torch.uint8 for im in images) else torch.float32
        batch = torch.zeros(len(images), 3, H, W, dtype=dtype, device=self.device)
        for i, im in enumerate(images):
            batch[i, :, : sizes[i][0], : sizes[i][1]] = im.to(device=self.device, dtype=dtype, non_blocking=True)
        return batch, sizes

    @torch.no_grad()
    def forward(self, batched_inputs: Sequence[dict], do_postprocess: bool = True):
        if self.training:
            raise NotImplementedError("training is outside the hot-path scope")
        return self._collect([self._launch(batched_inputs, do_postprocess)])[0]

    @torch.no_grad()
    def _launch(self, batched_inputs: Sequence[dict], do_postprocess: bool = True):
        """Enqueue one batch on the current stream and return its device-side result WITHOUT synchronising: callers
        that run several batches back to back (test-time augmentation: 9 batches per image) launch them all and pay
        one host sync in `_collect` instead of one per batch."""
        batch, sizes = self.preprocess_image(batched_inputs)
        eng = self._get_engine((int(batch.shape[0]), int(batch.shape[2]), int(batch.shape[3])))
        out_sizes = [
            (int(inp.get("height", s[0])), int(inp.get("width", s[1]))) for inp, s in zip(batched_inputs, sizes)
        ]
        with torch.cuda.device(self.device):
            if self.use_cuda_graphs:
                # One captured step per engine (= per batch shape), bound to a private input buffer, these image /
                # output sizes and a private result record. A caller that repeats shapes and sizes -- TTA runs the
                # same 9 scales for every image of a dataset -- replays it: one cudaGraphLaunch instead of ~200
                # launches (2.5 ms of host time per step at batch 1). Anything else re-captures or runs eagerly.
                key = (batch.dtype, tuple(sizes), tuple(out_sizes), bool(do_postprocess), self._weights_version)
                st = getattr(eng, "_graph_state", None)
                if st is not None and st["busy"]:
                    st = None  # its result record has not been collected yet: this batch runs eagerly
                elif st is not None and st["key"] == key:
                    st["buf"].copy_(batch)
                    eng.replay()
                    st["busy"] = True
                    return st["wire"].dets, st["wire"].counts, out_sizes, st
                else:
                    from .engine import DetectionWire

                    buf = batch.clone()
                    wire = DetectionWire(batch.shape[0], self.spec.post_nms_topk + 64, self.device)
                    eng.capture(buf, sizes, out_sizes, do_postprocess, out=wire)  # its eager step serves this batch
                    st = dict(key=key, buf=buf, wire=wire, busy=True)
                    eng._graph_state = st
                    return wire.dets, wire.counts, out_sizes, st
            dets, counts = eng.detect(batch, sizes, out_sizes, do_postprocess=do_postprocess)
        return dets, counts, out_sizes, None

    def _collect(self, launched: Sequence[tuple]):
        """[(dets, counts, out_sizes)] from `_launch` -> per batch the reference's [{"instances": Instances}]; the counts
        of ALL batches come to the host in one copy (the one sync)."""
        if not launched:
            return []
        with torch.cuda.device(self.device):
            counts_h = torch.cat([c for _, c, _, _ in launched]).cpu().tolist()
        out, k = [], 0
        for dets, counts, out_sizes, graph_state in launched:
            cap = dets.shape[1]
            results = []
            # one contiguous copy per FIELD for the whole batch (seven launches per batch, not per image); an image's
            # fields are the leading rows of its slice
            f_boxes, f_corners = dets[:, :, 8:12].contiguous(), dets[:, :, 0:8].contiguous()
            f_scores, f_ctr = dets[:, :, 12].contiguous(), dets[:, :, 13].contiguous()
            f_cls, f_lvl = dets[:, :, 14].to(torch.int64), dets[:, :, 15].to(torch.int64)
            f_loc = dets[:, :, 16:18].contiguous()
            for i in range(dets.shape[0]):
                n = counts_h[k]
                k += 1
                if n > cap:
                    raise _capi.DafneError(f"{n} detections exceed the output capacity {cap} (score ties at the cut)")
                # detectron2's detector_postprocess runs inside ProposalNetwork.forward whatever do_postprocess says:
                # the boxes are scaled / clipped / filtered and image_size is the requested output size either way;
                # do_postprocess only gates the corner / location rescale (one_stage_detector.py:45-55, 78-98)
                inst = Instances(out_sizes[i])
                inst.pred_boxes = Boxes(f_boxes[i, :n])
                inst.pred_corners = f_corners[i, :n]
                inst.scores = f_scores[i, :n]
                inst.centerness = f_ctr[i, :n]
                inst.pred_classes = f_cls[i, :n]
                inst.locations = f_loc[i, :n]
                inst.fpn_levels = f_lvl[i, :n]
                results.append({"instances": inst})
            if graph_state is not None:
                graph_state["busy"] = False  # every field above is a copy: the record may be overwritten again
            out.append(results)
        return out

    def inference(self, batched_inputs, detected_instances=None, do_postprocess: bool = True):
        assert not self.training
        return self.forward(batched_inputs, do_postprocess)

    @staticmethod
    def _postprocess(processed_results, batched_inputs):
        """Rescale corners / locations to the requested output size (one_stage_detector.py:78-98), for callers that
        ran `forward(..., do_postprocess=False)` and rescale later (TTA)."""
        for res, inp in zip(processed_results, batched_inputs):
            key = "proposals" if "proposals" in res else "instances"
            oh, ow = inp["height"], inp["width"]
            ih, iw = inp["image"].shape[1:3]
            sx, sy = ow / iw, oh / ih
            r = res[key]
            r.pred_corners[:, 0::2] *= sx
            r.pred_corners[:, 1::2] *= sy
            r.locations[:, 0] *= sx
            r.locations[:, 1] *= sy
        return processed_results


def build_model(cfg) -> nn.Module:
    """detectron2.modeling.build_model: look the meta-architecture up by name and move it to cfg.MODEL.DEVICE."""
    model = META_ARCH_REGISTRY.get(cfg.MODEL.META_ARCHITECTURE)(cfg)
    model.to(torch.device(cfg.MODEL.DEVICE))
    return model


class DefaultPredictor:
    """detectron2.engine.DefaultPredictor for this model (call shape of tools/vis/feature_maps.py:171-194):
    predictor(bgr_hwc_uint8) -> {"instances": Instances} in the coordinates of the original image. Like detectron2's,
    it resizes with ResizeShortestEdge(INPUT.MIN_SIZE_TEST, INPUT.MAX_SIZE_TEST) first -- here on the device, with the
    Pillow-exact bilinear kernel (`resize=False` feeds the image at its own resolution)."""

    def __init__(self, cfg, state_dict: Optional[Dict[str, torch.Tensor]] = None, resize: bool = True):
        self.cfg = cfg
        self.model = build_model(cfg)
        if state_dict is not None:
            self.model.load_state_dict(state_dict)
        self.input_format = cfg.INPUT.FORMAT
        self.resize = resize
        self.min_size_test, self.max_size_test = int(cfg.INPUT.MIN_SIZE_TEST), int(cfg.INPUT.MAX_SIZE_TEST)

    def __call__(self, original_image: np.ndarray):
        from .tta import NoOpTransform, resize_shortest_edge_transform

        if self.input_format == "RGB":
            original_image = original_image[:, :, ::-1]
        h, w = original_image.shape[:2]
        image = torch.as_tensor(np.ascontiguousarray(original_image.transpose(2, 0, 1)))
        if self.resize and image.dtype == torch.uint8:
            t = resize_shortest_edge_transform(h, w, self.min_size_test, self.max_size_test)
            if not isinstance(t, NoOpTransform):
                image = resize_bilinear_u8(image.to(self.model.device), t.new_h, t.new_w)
        return self.model([{"image": image, "height": h, "width": w}])[0]
