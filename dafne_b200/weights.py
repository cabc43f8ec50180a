"""State-dict layout of the DAFNe inference model and the seeded synthetic-weight recipe used for benchmarks/tests.

Names follow the module tree the reference builds (detectron2 v0.5 ResNet/FPN under `backbone.`, the head under
`proposal_generator.dafne_head.`; dafne/modeling/backbone/fpn.py:26-27, dafne/modeling/dafne/dafne.py:209-258,310-347),
so a detectron2 checkpoint's `{"model": {...}}` dict loads by name. Published weights are not available offline
(README.md:48-53 are Google-Drive links); `synthetic_state_dict` produces random-init weights of the same architecture
that (i) keep fp16 activations in range and (ii) give the post-processing real work (thousands of overlapping,
elongated, rotated candidates) -- the reference's own default init yields no candidate above the 0.05 threshold.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Tuple

import torch

from .spec import ModelSpec

BU = "backbone.bottom_up."
HEAD = "proposal_generator.dafne_head."
TOWERS = ("cls_tower", "center_tower", "corners_tower")
STAGE_BLOCKS = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3)}


def _conv_bn(shapes, prefix, cout, cin, k):
    shapes[prefix + ".weight"] = (cout, cin, k, k)
    for s in ("weight", "bias", "running_mean", "running_var"):
        shapes[f"{prefix}.norm.{s}"] = (cout,)


def _conv_bias(shapes, prefix, cout, cin, k):
    shapes[prefix + ".weight"] = (cout, cin, k, k)
    shapes[prefix + ".bias"] = (cout,)


def state_dict_shapes(spec: ModelSpec) -> "OrderedDict[str, Tuple[int, ...]]":
    """Every tensor the inference path needs, in module order, with its shape in the reference's native layout."""
    shapes: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    _conv_bn(shapes, BU + "stem.conv1", 64, 3, 7)
    cin, mid, cout = 64, 64, 256
    for s, nblocks in zip(range(2, 6), STAGE_BLOCKS[spec.resnet_depth]):
        for b in range(nblocks):
            pre = f"{BU}res{s}.{b}"
            if b == 0:
                _conv_bn(shapes, pre + ".shortcut", cout, cin, 1)
            _conv_bn(shapes, pre + ".conv1", mid, cin, 1)
            _conv_bn(shapes, pre + ".conv2", mid, mid, 3)
            _conv_bn(shapes, pre + ".conv3", cout, mid, 1)
            cin = cout
        mid, cout = mid * 2, cout * 2
    for i, c in zip((3, 4, 5), (512, 1024, 2048)):
        _conv_bias(shapes, f"backbone.fpn_lateral{i}", 256, c, 1)
        _conv_bias(shapes, f"backbone.fpn_output{i}", 256, 256, 3)
    _conv_bias(shapes, "backbone.top_block.p6", 256, 256, 3)
    _conv_bias(shapes, "backbone.top_block.p7", 256, 256, 3)
    for t in TOWERS:
        for i in range(4):
            _conv_bias(shapes, f"{HEAD}{t}.{3 * i}", 256, 256, 3)
            shapes[f"{HEAD}{t}.{3 * i + 1}.weight"] = (256,)
            shapes[f"{HEAD}{t}.{3 * i + 1}.bias"] = (256,)
    _conv_bias(shapes, HEAD + "cls_logits", spec.num_classes, 256, 3)
    _conv_bias(shapes, HEAD + "ctrness", 1, 256, 3)
    _conv_bias(shapes, HEAD + "corners_pred", 8, 256, 3)
    _conv_bias(shapes, HEAD + "center_pred", 2, 256, 3)
    for l in range(5):
        shapes[f"{HEAD}scales.{l}.scale"] = (1,)
    return shapes


# Frozen constants of the synthetic recipe (calibrated once on 1024^2 uniform-noise images):
# the class bias puts roughly 1 % of the (location, class) scores of a level above the 0.05 threshold.
SYNTH_CLS_STD = 0.02
SYNTH_CLS_BIAS = -4.6  # reference default: -log((1 - 0.01) / 0.01) (dafne.py:283-285) -> no candidates at all
# (depth, THRESH_WITH_CTR) -> class bias giving roughly 0.5-2 % of (location, class) scores above the threshold on
# uniform-noise 1024^2 images (scripts/dev_calibrate.py, gpurun_out/calib.log of round 1).
SYNTH_CLS_BIAS_TABLE = {(50, False): -4.25, (50, True): -6.5, (101, False): -4.4, (101, True): -6.75}
SYNTH_BASE_QUAD = (-3.0, -1.0, 3.0, -1.0, 3.0, 1.0, -3.0, 1.0)  # stride units: 48x16 px at p3, 96x32 at p4, ...


def synthetic_state_dict(spec: ModelSpec, seed: int = 0, cls_bias: float | None = None,
                         base_quad=SYNTH_BASE_QUAD) -> Dict[str, torch.Tensor]:
    """Seeded random-init weights (CPU fp32) of the architecture `spec` describes."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd: Dict[str, torch.Tensor] = OrderedDict()
    shapes = state_dict_shapes(spec)

    def kaiming(shape):
        cout, _, kh, kw = shape
        std = math.sqrt(2.0 / (cout * kh * kw))  # kaiming_normal_(mode="fan_out", nonlinearity="relu")
        return torch.randn(shape, generator=g) * std

    for name, shape in shapes.items():
        if name.endswith(".norm.weight"):
            # last BN of every bottleneck is damped so the residual stream stays inside fp16 range over 33 blocks
            sd[name] = torch.full(shape, 0.25 if ".conv3." in name else 1.0)
        elif name.endswith(".norm.bias") or name.endswith(".norm.running_mean"):
            sd[name] = torch.zeros(shape)
        elif name.endswith(".norm.running_var"):
            sd[name] = torch.ones(shape)
        elif name.endswith(".scale"):
            sd[name] = torch.ones(shape)
        elif len(shape) == 4:
            if name == HEAD + "cls_logits.weight":
                sd[name] = torch.randn(shape, generator=g) * SYNTH_CLS_STD
            elif name == HEAD + "ctrness.weight":
                sd[name] = torch.randn(shape, generator=g) * 0.01
            elif name == HEAD + "corners_pred.weight":
                sd[name] = torch.randn(shape, generator=g) * 0.02
            elif name == HEAD + "center_pred.weight":
                sd[name] = torch.randn(shape, generator=g) * 0.02
            else:
                sd[name] = kaiming(shape)
        elif name == HEAD + "cls_logits.bias":
            auto = SYNTH_CLS_BIAS_TABLE[(spec.resnet_depth, bool(spec.thresh_with_ctr))]
            sd[name] = torch.full(shape, auto if cls_bias is None else cls_bias)
        elif name == HEAD + "corners_pred.bias":
            sd[name] = torch.tensor(base_quad, dtype=torch.float32)
        elif any(f"{t}." in name for t in TOWERS) and name.endswith(".weight"):
            sd[name] = torch.ones(shape)  # GroupNorm gamma
        else:
            sd[name] = torch.zeros(shape)  # conv biases, GroupNorm beta
    return sd
