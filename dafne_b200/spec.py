"""Flat description of the model the hot path runs, derived from a config (see include/dafne_b200.h)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Tuple

from . import _capi


@dataclass
class ModelSpec:
    resnet_depth: int = 50
    num_classes: int = 15
    sort_corners: bool = True
    thresh_with_ctr: bool = False
    pre_nms_topk: int = 2000
    post_nms_topk: int = 1000
    score_thresh: float = 0.05
    nms_thresh: float = 0.1
    fpn_strides: Tuple[int, ...] = (8, 16, 32, 64, 128)
    pixel_mean: Tuple[float, ...] = (123.675, 116.28, 103.53)
    pixel_std: Tuple[float, ...] = (1.0, 1.0, 1.0)
    vehicle_merge: bool = True  # reference behaviour for every dataset (dafne/modeling/nms/nms.py:77-79)
    size_divisibility: int = 32
    unsupported: List[str] = field(default_factory=list)

    @staticmethod
    def from_cfg(cfg) -> "ModelSpec":
        """Read the keys the reference modules read (dafne.py:168-285, dafne_outputs.py:123-193, fpn.py:58-91) and
        refuse configurations outside the hot-path scope instead of silently computing something else."""
        m, d = cfg.MODEL, cfg.MODEL.DAFNE
        problems = []

        def need(cond, msg):
            if not cond:
                problems.append(msg)

        need(m.META_ARCHITECTURE == "OneStageDetector", f"META_ARCHITECTURE={m.META_ARCHITECTURE}")
        need(m.BACKBONE.NAME == "build_dafne_resnet_fpn_backbone", f"BACKBONE.NAME={m.BACKBONE.NAME}")
        need(m.PROPOSAL_GENERATOR.NAME == "DAFNe", f"PROPOSAL_GENERATOR.NAME={m.PROPOSAL_GENERATOR.NAME}")
        need(not m.BACKBONE.get("ANTI_ALIAS", False), "BACKBONE.ANTI_ALIAS")
        need(m.RESNETS.get("DEFORM_INTERVAL", 1) <= 1, "RESNETS.DEFORM_INTERVAL > 1")
        need(not any(m.RESNETS.get("DEFORM_ON_PER_STAGE", [False])), "RESNETS.DEFORM_ON_PER_STAGE")
        need(m.RESNETS.DEPTH in (50, 101), f"RESNETS.DEPTH={m.RESNETS.DEPTH}")
        need(m.RESNETS.NORM == "FrozenBN", f"RESNETS.NORM={m.RESNETS.NORM}")
        need(m.RESNETS.STRIDE_IN_1X1, "RESNETS.STRIDE_IN_1X1=False")
        need(m.RESNETS.NUM_GROUPS == 1 and m.RESNETS.WIDTH_PER_GROUP == 64, "ResNeXt widths")
        need(m.RESNETS.RES5_DILATION == 1, "RES5_DILATION != 1")
        need(list(m.FPN.IN_FEATURES) == ["res3", "res4", "res5"], f"FPN.IN_FEATURES={m.FPN.IN_FEATURES}")
        need(m.FPN.OUT_CHANNELS == 256 and m.FPN.NORM == "" and m.FPN.FUSE_TYPE == "sum", "FPN variant")
        need(d.TOP_LEVELS == 2, f"DAFNE.TOP_LEVELS={d.TOP_LEVELS}")
        need(d.NORM == "GN", f"DAFNE.NORM={d.NORM}")
        need(d.CORNER_PREDICTION == "center-to-corner", f"DAFNE.CORNER_PREDICTION={d.CORNER_PREDICTION}")
        need(d.CORNER_TOWER_ON_CENTER_TOWER and not d.MERGE_CORNER_CENTER_PRED, "corner/center tower wiring")
        need(d.CENTERNESS != "none" and d.CTR_ON_REG, "centerness wiring")
        need(d.USE_SCALE and d.ENABLE_FPN_STRIDE_NORM, "USE_SCALE / ENABLE_FPN_STRIDE_NORM")
        need(not d.USE_DEFORMABLE, "DAFNE.USE_DEFORMABLE")
        need(d.NUM_CLS_CONVS == 4 and d.NUM_BOX_CONVS == 4 and d.NUM_SHARE_CONVS == 0, "tower depths")
        need(list(d.FPN_STRIDES) == [8, 16, 32, 64, 128], f"FPN_STRIDES={d.FPN_STRIDES}")
        need(1 <= d.NUM_CLASSES <= 32, f"NUM_CLASSES={d.NUM_CLASSES}")
        need(m.TOP_MODULE.NAME in ("", None), f"TOP_MODULE.NAME={m.TOP_MODULE.NAME}")
        need(d.NMS_TH > 0, "NMS_TH <= 0")
        if problems:
            raise NotImplementedError(
                "configuration outside the B200 hot-path scope (SURVEY.md section 8): " + "; ".join(problems)
            )
        return ModelSpec(
            resnet_depth=int(m.RESNETS.DEPTH),
            num_classes=int(d.NUM_CLASSES),
            sort_corners=bool(d.SORT_CORNERS),
            thresh_with_ctr=bool(d.THRESH_WITH_CTR),
            pre_nms_topk=int(d.PRE_NMS_TOPK_TEST),
            post_nms_topk=int(d.POST_NMS_TOPK_TEST),
            score_thresh=float(d.INFERENCE_TH_TEST),
            nms_thresh=float(d.NMS_TH),
            fpn_strides=tuple(int(s) for s in d.FPN_STRIDES),
            pixel_mean=tuple(float(v) for v in m.PIXEL_MEAN),
            pixel_std=tuple(float(v) for v in m.PIXEL_STD),
        )

    def to_c(self) -> "_capi.ModelSpecC":
        c = _capi.ModelSpecC()
        c.resnet_depth = self.resnet_depth
        c.num_classes = self.num_classes
        c.sort_corners = int(self.sort_corners)
        c.thresh_with_ctr = int(self.thresh_with_ctr)
        c.pre_nms_topk = self.pre_nms_topk
        c.post_nms_topk = self.post_nms_topk
        c.score_thresh = self.score_thresh
        c.nms_thresh = self.nms_thresh
        c.num_levels = len(self.fpn_strides)
        for i, s in enumerate(self.fpn_strides):
            c.fpn_strides[i] = s
        for i in range(3):
            c.pixel_mean[i] = self.pixel_mean[i]
            c.pixel_std[i] = self.pixel_std[i]
        c.vehicle_merge = int(self.vehicle_merge)
        return c

    def level_sizes(self, H: int, W: int):
        """(H_l, W_l) of p3..p7 for a padded H x W input (H, W multiples of 32)."""
        out = [(H // 8, W // 8), (H // 16, W // 16), (H // 32, W // 32)]
        for _ in range(2):
            h, w = out[-1]
            out.append(((h - 1) // 2 + 1, (w - 1) // 2 + 1))
        return out
