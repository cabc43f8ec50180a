"""Test-time augmentation around the B200 hot path: the reference's dafne/modeling/tta.py with the same
class names, constructor arguments and call shape.

  DotaDatasetMapperTTA(cfg)(dataset_dict) -> [augmented dataset dicts with "transforms"]      tta.py:29-135
  OneStageRCNNWithTTA(cfg, model, tta_mapper=None, batch_size=3)(batched_inputs)               tta.py:138-268

What changes underneath: every augmented copy runs through `model.inference(..., do_postprocess=False)` = the CUDA hot
path; the corners are mapped back to the original image ON THE DEVICE (the reference copies them to the host, applies
`tfm.inverse().apply_coords` in numpy and copies them back, tta.py:251-259); the union (up to 27 copies x 1000 boxes)
goes through the same device polygon NMS (`select_over_all_levels`, tta.py:264-268).

The transforms are the small subset of detectron2 v0.5 / fvcore the mapper uses (not installable offline; restated from
their published behaviour): ResizeShortestEdge / Resize -> ResizeTransform (PIL bilinear for uint8 HWC images),
RandomFlip(prob=1) -> HFlipTransform / VFlipTransform, TransformList with inverse(). The augmented images themselves are
produced on the host exactly like the reference's mapper does (input side, outside the hot path); rotation TTA
(TEST.AUG.ROTATION_ANGLES, empty in every pre-trained config) is not implemented and raises.
"""
from __future__ import annotations

import copy
from itertools import count
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch
from torch import nn

from .structures import Instances


# ------------------------------------------------------------------------------------------------ transforms
class Transform:
    def apply_image(self, img: np.ndarray) -> np.ndarray:
        raise NotImplementedError

    def apply_coords(self, coords: np.ndarray) -> np.ndarray:
        raise NotImplementedError

    def apply_coords_device(self, coords: torch.Tensor) -> torch.Tensor:
        """Same arithmetic as apply_coords (float32 coordinates, python scalars), on a [n, 2] device tensor."""
        raise NotImplementedError

    def inverse(self) -> "Transform":
        raise NotImplementedError

    def __add__(self, other):
        a = self.transforms if isinstance(self, TransformList) else [self]
        b = other.transforms if isinstance(other, TransformList) else [other]
        return TransformList(a + b)


class NoOpTransform(Transform):
    def apply_image(self, img):
        return img

    def apply_image_device(self, img):
        return img

    def apply_coords(self, coords):
        return coords

    def apply_coords_device(self, coords):
        return coords

    def inverse(self):
        return self


class HFlipTransform(Transform):
    def __init__(self, width: int):
        self.width = width

    def apply_image(self, img):
        return np.flip(img, axis=1)

    def apply_image_device(self, img):
        return img.flip(-1)

    def apply_coords(self, coords):
        coords[:, 0] = self.width - coords[:, 0]
        return coords

    def apply_coords_device(self, coords):
        coords[:, 0] = self.width - coords[:, 0]
        return coords

    def inverse(self):
        return self


class VFlipTransform(Transform):
    def __init__(self, height: int):
        self.height = height

    def apply_image(self, img):
        return np.flip(img, axis=0)

    def apply_image_device(self, img):
        return img.flip(-2)

    def apply_coords(self, coords):
        coords[:, 1] = self.height - coords[:, 1]
        return coords

    def apply_coords_device(self, coords):
        coords[:, 1] = self.height - coords[:, 1]
        return coords

    def inverse(self):
        return self


class ResizeTransform(Transform):
    """detectron2 ResizeTransform: PIL bilinear for uint8 images, coordinates scaled by new / old."""

    def __init__(self, h: int, w: int, new_h: int, new_w: int, interp=None):
        self.h, self.w, self.new_h, self.new_w, self.interp = h, w, new_h, new_w, interp

    def apply_image(self, img):
        assert img.shape[:2] == (self.h, self.w), (img.shape, self.h, self.w)
        if img.dtype != np.uint8:
            raise NotImplementedError("ResizeTransform: uint8 HWC images only (what the TTA mapper feeds it)")
        from PIL import Image

        pil = Image.fromarray(img if img.shape[2] != 1 else img[:, :, 0])
        pil = pil.resize((self.new_w, self.new_h), Image.BILINEAR if self.interp is None else self.interp)
        ret = np.asarray(pil)
        return ret if ret.ndim == 3 else ret[:, :, None]

    def apply_image_device(self, img: torch.Tensor) -> torch.Tensor:
        """[C, H, W] uint8 CUDA tensor: the same Pillow bilinear resize, on the device (dafne_resize_bilinear_u8)."""
        from .modeling import resize_bilinear_u8

        assert tuple(img.shape[-2:]) == (self.h, self.w), (tuple(img.shape), self.h, self.w)
        return resize_bilinear_u8(img, self.new_h, self.new_w)

    def apply_coords(self, coords):
        coords[:, 0] = coords[:, 0] * (self.new_w * 1.0 / self.w)
        coords[:, 1] = coords[:, 1] * (self.new_h * 1.0 / self.h)
        return coords

    def apply_coords_device(self, coords):
        coords[:, 0] = coords[:, 0] * (self.new_w * 1.0 / self.w)
        coords[:, 1] = coords[:, 1] * (self.new_h * 1.0 / self.h)
        return coords

    def inverse(self):
        return ResizeTransform(self.new_h, self.new_w, self.h, self.w, self.interp)


class TransformList(Transform):
    def __init__(self, transforms: Sequence[Transform]):
        self.transforms = list(transforms)

    def apply_image(self, img):
        for t in self.transforms:
            img = t.apply_image(img)
        return img

    def apply_coords(self, coords):
        for t in self.transforms:
            coords = t.apply_coords(coords)
        return coords

    def apply_coords_device(self, coords):
        for t in self.transforms:
            coords = t.apply_coords_device(coords)
        return coords

    def inverse(self):
        return TransformList([t.inverse() for t in self.transforms[::-1]])


def resize_shortest_edge_transform(h: int, w: int, size: int, max_size: int) -> Transform:
    """detectron2 v0.5 ResizeShortestEdge.get_transform with a fixed short-edge length."""
    if size == 0:
        return NoOpTransform()
    scale = size * 1.0 / min(h, w)
    if h < w:
        newh, neww = size, scale * w
    else:
        newh, neww = scale * h, size
    if max(newh, neww) > max_size:
        scale = max_size * 1.0 / max(newh, neww)
        newh = newh * scale
        neww = neww * scale
    return ResizeTransform(h, w, int(newh + 0.5), int(neww + 0.5))


# ------------------------------------------------------------------------------------------------ mapper (tta.py:29-135)
class DotaDatasetMapperTTA:
    """`device` (not in the reference): build the copies on that CUDA device -- one H2D copy of the image, the
    Pillow-exact resize kernel and flips there -- instead of on the host; the copies are bit-identical either way."""

    def __init__(self, cfg, device=None):
        self.device = torch.device(device) if device is not None else None
        self.min_sizes = list(cfg.TEST.AUG.MIN_SIZES)
        self.max_size = cfg.TEST.AUG.MAX_SIZE
        self.resize_type = cfg.INPUT.RESIZE_TYPE
        self.vflip = cfg.TEST.AUG.VFLIP
        self.hflip = cfg.TEST.AUG.HFLIP
        self.rotation_angles = list(cfg.TEST.AUG.ROTATION_ANGLES)
        self.image_format = cfg.INPUT.FORMAT
        self.cfg = cfg
        if len(self.rotation_angles) != 0:
            raise NotImplementedError("TEST.AUG.ROTATION_ANGLES: rotation TTA is outside the built scope "
                                      "(empty in every pre-trained config)")

    def __call__(self, dataset_dict):
        numpy_image = dataset_dict["image"].permute(1, 2, 0).cpu().numpy()
        shape = numpy_image.shape
        orig_shape = (dataset_dict["height"], dataset_dict["width"])
        if shape[:2] != orig_shape:
            pre_tfm: Transform = ResizeTransform(orig_shape[0], orig_shape[1], shape[0], shape[1])
        else:
            pre_tfm = NoOpTransform()
        candidates: List[List[Callable[[np.ndarray], Transform]]] = []
        for min_size in self.min_sizes:
            if self.resize_type == "shortest-edge":
                def resize(img, s=min_size):
                    return resize_shortest_edge_transform(img.shape[0], img.shape[1], s, self.max_size)
            elif self.resize_type == "both":
                h_test, w_test = self.cfg.INPUT.RESIZE_HEIGHT_TEST, self.cfg.INPUT.RESIZE_WIDTH_TEST
                new_h, new_w = int(h_test * (min_size / w_test)), min_size

                def resize(img, nh=new_h, nw=new_w):
                    return ResizeTransform(img.shape[0], img.shape[1], nh, nw)
            else:
                raise RuntimeError(f"Invalid resize-type: {self.resize_type}")
            candidates.append([resize])
            if self.hflip:
                candidates.append([resize, lambda img: HFlipTransform(img.shape[1])])
            if self.vflip:
                candidates.append([resize, lambda img: VFlipTransform(img.shape[0])])
        ret = []
        on_device = self.device is not None and dataset_dict["image"].dtype == torch.uint8
        dev_image = dataset_dict["image"].to(self.device) if on_device else None
        for aug in candidates:
            img = np.copy(numpy_image) if not on_device else None
            cur = dev_image
            shape_hw = numpy_image.shape[:2]
            tfms = []
            for make in aug:  # detectron2 apply_augmentations: each transform is built from the current image
                t = make(np.empty(shape_hw + (0,), np.uint8) if on_device else img)
                if on_device:
                    cur = t.apply_image_device(cur)
                    shape_hw = tuple(cur.shape[-2:])
                else:
                    img = t.apply_image(img)
                tfms.append(t)
            dic = copy.copy(dataset_dict)
            dic["transforms"] = pre_tfm + TransformList(tfms)
            dic["image"] = cur.contiguous() if on_device else torch.from_numpy(np.ascontiguousarray(img.transpose(2, 0, 1)))
            ret.append(dic)
        return ret


# ------------------------------------------------------------------------------------------------ model (tta.py:138-268)
class OneStageRCNNWithTTA(nn.Module):
    def __init__(self, cfg, model, tta_mapper: Optional[Callable] = None, batch_size: int = 3,
                 use_cuda_graphs: bool = True, cross_image_batch: int = 12):
        """`use_cuda_graphs` (not in the reference): replay each augmented shape as one captured CUDA graph -- the copies
        of every image of a dataset repeat the same shapes and sizes, and at batch 3 a step is bound by the host's
        ~200 launches, not by the GPU.
        `cross_image_batch` (not in the reference): when ONE call holds several images, copies of the same shape from
        different images run as one batch of up to this many copies instead of `batch_size` copies of one image (the
        reference loops over the images, tta.py:178-196). A copy's detections do not depend on what else is in its batch
        (per-image GroupNorm statistics, per-image NMS), so the results are those of the per-image loop; batches of 12
        instead of 3 simply fill the GPU. <= batch_size: the reference's batches."""
        super().__init__()
        from .modeling import OneStageDetector

        assert isinstance(model, OneStageDetector), \
            "TTA is only supported on OneStageDetector. Got a model of type {}".format(type(model))
        self.cfg = cfg
        self.model = model
        self.tta_mapper = DotaDatasetMapperTTA(cfg) if tta_mapper is None else tta_mapper
        self.batch_size = batch_size
        self.use_cuda_graphs = use_cuda_graphs
        self.cross_image_batch = cross_image_batch

    def _batch_inference(self, batched_inputs, detected_instances=None):
        outputs = []
        inputs = []
        for idx, inp in zip(count(), batched_inputs):
            inputs.append(inp)
            if len(inputs) == self.batch_size or idx == len(batched_inputs) - 1:
                outputs.extend(self.model.inference(inputs, None, do_postprocess=False))
                inputs = []
        return outputs

    def __call__(self, batched_inputs):
        def _maybe_read_image(dataset_dict):
            ret = copy.copy(dataset_dict)
            if "image" not in ret:
                raise NotImplementedError("file_name inputs: reading images from disk is outside the built scope")
            if "height" not in ret and "width" not in ret:
                ret["height"] = ret["image"].shape[1]
                ret["width"] = ret["image"].shape[2]
            return ret

        inputs = [_maybe_read_image(x) for x in batched_inputs]
        if len(inputs) > 1 and self.cross_image_batch > self.batch_size:
            return self._inference_images_batched(inputs)
        return [self._inference_one_image(x) for x in inputs]

    def _inference_images_batched(self, inputs):
        """All images of the call at once: their copies are grouped by shape across images (see `cross_image_batch`),
        every group is enqueued, ONE host sync collects all of them, then each image's copies are merged as in
        `_inference_one_image`."""
        per_image = [self._get_augmented_inputs(x) for x in inputs]
        groups = {}
        for k, (aug, _) in enumerate(per_image):
            for j, a in enumerate(aug):
                key = (tuple(a["image"].shape), a["image"].dtype, a.get("height"), a.get("width"))
                groups.setdefault(key, []).append((k, j, a))
        launched, owners = [], []
        saved = self.model.use_cuda_graphs
        self.model.use_cuda_graphs = self.use_cuda_graphs
        try:
            for items in groups.values():
                for s0 in range(0, len(items), self.cross_image_batch):
                    chunk = items[s0 : s0 + self.cross_image_batch]
                    launched.append(self.model._launch([a for _, _, a in chunk], do_postprocess=False))
                    owners.append([(k, j) for k, j, _ in chunk])
        finally:
            self.model.use_cuda_graphs = saved
        outputs = [[None] * len(aug) for aug, _ in per_image]
        for res, own in zip(self.model._collect(launched), owners):
            for o, (k, j) in zip(res, own):
                outputs[k][j] = o
        # the union NMS of all images side by side (select_over_all_levels overlaps the images of its list)
        merged = self.model.proposal_generator.dafne_outputs.select_over_all_levels(
            [self._to_original_frame(outputs[k], per_image[k][1]) for k in range(len(inputs))])
        return [{"instances": m} for m in merged]

    def _inference_one_image(self, input):
        augmented_inputs, tfms = self._get_augmented_inputs(input)
        instances = self._get_augmented_corners(augmented_inputs, tfms)
        return {"instances": self._merge_detections(instances)}

    def _get_augmented_inputs(self, input):
        augmented_inputs = self.tta_mapper(input)
        tfms = [x.pop("transforms") for x in augmented_inputs]
        return augmented_inputs, tfms

    def _batch_inference_deferred(self, batched_inputs):
        """`_batch_inference` without its host sync per batch: every batch of `batch_size` copies is enqueued first (one
        engine per copy shape, so nothing waits on anything but the stream), then the result sizes of all of them come
        back in one copy. Same batches, same kernels, same results as `_batch_inference`."""
        launched, inputs = [], []
        saved = self.model.use_cuda_graphs
        self.model.use_cuda_graphs = self.use_cuda_graphs
        try:
            for idx, inp in zip(count(), batched_inputs):
                inputs.append(inp)
                if len(inputs) == self.batch_size or idx == len(batched_inputs) - 1:
                    launched.append(self.model._launch(inputs, do_postprocess=False))
                    inputs = []
        finally:
            self.model.use_cuda_graphs = saved
        return [o for batch in self.model._collect(launched) for o in batch]

    def _get_augmented_corners(self, augmented_inputs, tfms):
        return self._to_original_frame(self._batch_inference_deferred(augmented_inputs), tfms)

    def _to_original_frame(self, outputs, tfms):
        instances_list = []
        for output, tfm in zip(outputs, tfms):
            instances = output["instances"]
            pred_corners = instances.pred_corners
            N, C = pred_corners.shape
            assert C == 8
            inv = tfm.inverse()
            if isinstance(inv, Transform):  # device: same float32 arithmetic as apply_coords, no host round trip
                original = inv.apply_coords_device(pred_corners.reshape(-1, 2).clone()).reshape(N, C)
            else:  # a foreign (e.g. detectron2) transform object: the reference's host path
                original = torch.from_numpy(inv.apply_coords(pred_corners.reshape(-1, 2).cpu().numpy()).reshape(N, C))
                original = original.to(pred_corners.device, dtype=pred_corners.dtype)
            inst = Instances(instances.image_size)
            inst.scores = instances.scores
            inst.centerness = instances.centerness
            inst.pred_corners = original
            inst.pred_classes = instances.pred_classes
            instances_list.append(inst)
        return Instances.cat(instances_list)

    def _merge_detections(self, instances):
        return self.model.proposal_generator.dafne_outputs.select_over_all_levels([instances])[0]
