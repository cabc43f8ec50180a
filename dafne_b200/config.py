"""Tolerant yacs-style config for the inference path.

Mirrors `dafne.config.get_cfg()` (reference dafne/config/config.py:4-13 on top of dafne/config/defaults.py:1-151) for the
keys the hot path reads, and loads the reference's YAML files unchanged: `_BASE_` chains
(configs/dota-1.0/1024.yaml:1), python literals in strings ("(1024,)"), YAML anchors, and the full detectron2 default
dumps of configs/pre-trained/*.yaml including keys this build does not know (they are kept, not rejected).
Neither detectron2 nor yacs is a dependency.
"""
from __future__ import annotations

import ast
import copy
import os
from typing import Any

import yaml

BASE_KEY = "_BASE_"


class CfgNode(dict):
    """dict with attribute access, recursive merge and `KEY VALUE` overrides (the subset of yacs the callers use)."""

    def __init__(self, init: dict | None = None):
        super().__init__()
        for k, v in (init or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, name: str) -> Any:
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e

    def __setattr__(self, name: str, value: Any) -> None:
        self[name] = value

    def clone(self) -> "CfgNode":
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        return CfgNode({k: copy.deepcopy(v, memo) for k, v in self.items()})

    # -- merging ---------------------------------------------------------------------------------
    def merge_from_other_cfg(self, other: dict) -> None:
        _merge(self, other)

    def merge_from_file(self, path: str) -> None:
        _merge(self, load_yaml_with_base(path))

    def merge_from_list(self, opts: list) -> None:
        if len(opts) % 2:
            raise ValueError("override list must be KEY VALUE pairs")
        for key, val in zip(opts[0::2], opts[1::2]):
            node = self
            parts = key.split(".")
            for p in parts[:-1]:
                if p not in node or not isinstance(node[p], dict):
                    node[p] = CfgNode()
                node = node[p]
            node[parts[-1]] = _coerce(val, node.get(parts[-1]))

    def freeze(self) -> None:  # kept for call-compatibility; configs are plain data here
        pass

    def defrost(self) -> None:
        pass

    def dump(self) -> str:
        return yaml.safe_dump(_to_plain(self), default_flow_style=None)


def _to_plain(node):
    if isinstance(node, dict):
        return {k: _to_plain(v) for k, v in node.items()}
    if isinstance(node, tuple):
        return list(node)
    return node


def _decode(value: Any) -> Any:
    """yacs decodes strings with literal_eval: "(1024,)" -> tuple, "1e-3" -> float; plain words stay strings."""
    if isinstance(value, dict):
        return CfgNode({k: _decode(v) for k, v in value.items()})
    if isinstance(value, str):
        try:
            return ast.literal_eval(value)
        except (ValueError, SyntaxError):
            return value
    return value


def _coerce(new: Any, old: Any) -> Any:
    new = _decode(new)
    if old is None or isinstance(old, dict):
        return new
    if isinstance(old, bool) and isinstance(new, (int, bool)):
        return bool(new)
    if isinstance(old, float) and isinstance(new, int) and not isinstance(new, bool):
        return float(new)
    if isinstance(old, tuple) and isinstance(new, list):
        return tuple(new)
    if isinstance(old, list) and isinstance(new, tuple):
        return list(new)
    return new


def _merge(dst: CfgNode, src: dict) -> None:
    for k, v in src.items():
        if k == BASE_KEY:
            continue
        if isinstance(v, dict):
            if k not in dst or not isinstance(dst[k], dict):
                dst[k] = CfgNode()
            _merge(dst[k], v)
        else:
            dst[k] = _coerce(v, dst.get(k))


def load_yaml_with_base(path: str) -> dict:
    """Load a YAML file, resolving `_BASE_` recursively (relative to the including file), child overriding base."""
    with open(path, "r") as f:
        cfg = yaml.safe_load(f) or {}  # safe_load resolves anchors / aliases
    if BASE_KEY in cfg:
        base_path = cfg.pop(BASE_KEY)
        if base_path.startswith("~"):
            base_path = os.path.expanduser(base_path)
        if not os.path.isabs(base_path) and "://" not in base_path:
            base_path = os.path.join(os.path.dirname(path), base_path)
        base = load_yaml_with_base(base_path)
        _merge_plain(base, cfg)
        return base
    return cfg


def _merge_plain(dst: dict, src: dict) -> None:
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge_plain(dst[k], v)
        else:
            dst[k] = v


# Defaults for every key the inference path reads. Values: detectron2 v0.5 defaults overridden by
# dafne/config/defaults.py (cited per block).
_DEFAULTS = {
    "VERSION": 2,
    "MODEL": {
        "META_ARCHITECTURE": "OneStageDetector",  # defaults.py:136
        "DEVICE": "cuda",
        "WEIGHTS": "",
        "PIXEL_MEAN": [103.530, 116.280, 123.675],  # detectron2 default (BGR); the DAFNe YAMLs override it
        "PIXEL_STD": [1.0, 1.0, 1.0],
        "MOBILENET": False,  # defaults.py:22
        "BACKBONE": {"NAME": "build_dafne_resnet_fpn_backbone", "FREEZE_AT": 2, "ANTI_ALIAS": False},  # :137, :23
        "RESNETS": {
            "DEPTH": 50,
            "OUT_FEATURES": ["res3", "res4", "res5"],  # defaults.py:138
            "NUM_GROUPS": 1,
            "NORM": "FrozenBN",
            "WIDTH_PER_GROUP": 64,
            "STRIDE_IN_1X1": True,
            "RES5_DILATION": 1,
            "RES2_OUT_CHANNELS": 256,
            "STEM_OUT_CHANNELS": 64,
            "DEFORM_ON_PER_STAGE": [False, False, False, False],
            "DEFORM_MODULATED": False,
            "DEFORM_NUM_GROUPS": 1,
            "DEFORM_INTERVAL": 1,  # defaults.py:24
        },
        "FPN": {"IN_FEATURES": ["res3", "res4", "res5"], "OUT_CHANNELS": 256, "NORM": "", "FUSE_TYPE": "sum"},
        "PROPOSAL_GENERATOR": {"NAME": "DAFNe", "MIN_SIZE": 0},  # defaults.py:140
        "TOP_MODULE": {"NAME": "", "DIM": 16},  # defaults.py:33-35
        "DAFNE": {  # defaults.py:40-108
            "NUM_CLASSES": 15,
            "IN_FEATURES": ["p3", "p4", "p5", "p6", "p7"],
            "FPN_STRIDES": [8, 16, 32, 64, 128],
            "PRIOR_PROB": 0.01,
            "INFERENCE_TH_TRAIN": 0.05,
            "INFERENCE_TH_TEST": 0.05,
            "NMS_TH": 0.1,
            "PRE_NMS_TOPK_TRAIN": 2000,
            "PRE_NMS_TOPK_TEST": 2000,
            "POST_NMS_TOPK_TRAIN": 1000,
            "POST_NMS_TOPK_TEST": 1000,
            "TOP_LEVELS": 2,
            "NORM": "GN",
            "USE_SCALE": True,
            "SORT_CORNERS": True,
            "SORT_CORNERS_DATALOADER": True,
            "CENTERNESS": "oriented",
            "CENTERNESS_ALPHA": 5,
            "CENTERNESS_USE_IN_SCORE": True,
            "CORNER_PREDICTION": "center-to-corner",
            "CORNER_TOWER_ON_CENTER_TOWER": True,
            "MERGE_CORNER_CENTER_PRED": False,
            "ENABLE_FPN_STRIDE_NORM": True,
            "THRESH_WITH_CTR": False,
            "CTR_ON_REG": True,
            "USE_RELU": True,
            "USE_DEFORMABLE": False,
            "NUM_CLS_CONVS": 4,
            "NUM_BOX_CONVS": 4,
            "NUM_SHARE_CONVS": 0,
            "YIELD_PROPOSAL": False,
        },
    },
    "INPUT": {
        "FORMAT": "BGR",
        "MIN_SIZE_TEST": 800,
        "MAX_SIZE_TEST": 1333,
        "RESIZE_TYPE": "shortest-edge",  # defaults.py:123
    },
    "TEST": {
        "DETECTIONS_PER_IMAGE": 100,
        "AUG": {
            "ENABLED": False,
            "MIN_SIZES": (400, 500, 600, 700, 800, 900, 1000, 1100, 1200),
            "MAX_SIZE": 4000,
            "FLIP": True,
            "VFLIP": True,  # defaults.py:112-116
            "HFLIP": True,
            "ROTATION_ANGLES": (),
        },
    },
    "DATASETS": {"TRAIN": (), "TEST": ()},
}


def get_cfg() -> CfgNode:
    """A fresh config with the defaults of the inference path (the counterpart of dafne.config.get_cfg)."""
    return CfgNode(copy.deepcopy(_DEFAULTS))
