"""ORACLE (test infrastructure only -- the product path never imports this package).

CPU restatement, with the torch ops detectron2 itself would call, of DAFNe's dense forward:

  normalise + pad            dafne/modeling/one_stage_detector.py:100-107  (ImageList.from_tensors: zero pad AFTER
                             normalising, size_divisibility 32)
  ResNet-50/101 bottom-up    detectron2 v0.5 build_resnet_backbone via dafne/modeling/backbone/fpn.py:72
                             (BasicStem 7x7 s2 + FrozenBN + ReLU + maxpool 3x3 s2; BottleneckBlock with the stride on
                             the 1x1 conv1, FrozenBN eps 1e-5 = F.batch_norm in eval mode, out += shortcut; relu)
  FPN + LastLevelP6P7        detectron2 v0.5 FPN.forward via fpn.py:83-90; fpn.py:16-37
  DAFNeHead (GN towers)      dafne/modeling/dafne/dafne.py:287-348 (towers), :350-369,388-414,462-471 (forward)

detectron2 is an un-vendored dependency (Dockerfile:21, v0.5); its semantics above are restated from its published
source and cannot be executed here, so this part of the oracle is "parity unpinned" by reference tests (the reference
has none). Two precisions:
  mode="fp32"  the reference's arithmetic (what the PyTorch path computes);
  mode="o16"   quantisation-matched: identical graph, but weights and every stored activation are rounded to fp16
               exactly where the CUDA path stores fp16 (fp32 accumulate) -- the yardstick for the tensor-core kernels.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

BU = "backbone.bottom_up."
HEAD = "proposal_generator.dafne_head."
STAGE_BLOCKS = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3)}


def preprocess(images: Sequence[torch.Tensor], pixel_mean, pixel_std, size_divisibility: int = 32):
    """list of CHW (uint8 / float) -> normalised, zero-padded N x 3 x H x W float32 and the un-padded sizes."""
    mean = torch.tensor(pixel_mean, dtype=torch.float32).view(3, 1, 1)
    std = torch.tensor(pixel_std, dtype=torch.float32).view(3, 1, 1)
    normed = [(x.to(torch.float32) - mean) / std for x in images]
    sizes = [(int(x.shape[1]), int(x.shape[2])) for x in images]
    H = max(s[0] for s in sizes)
    W = max(s[1] for s in sizes)
    d = size_divisibility
    H, W = (H + d - 1) // d * d, (W + d - 1) // d * d
    batch = torch.zeros(len(images), 3, H, W, dtype=torch.float32)
    for i, x in enumerate(normed):
        batch[i, :, : x.shape[1], : x.shape[2]] = x
    return batch, sizes


class _Net:
    def __init__(self, sd: Dict[str, torch.Tensor], mode: str):
        assert mode in ("fp32", "o16")
        self.sd = sd
        self.q = mode == "o16"

    def rnd(self, x):  # the fp16 store the CUDA path performs
        return x.half().float() if self.q else x

    def w(self, name):
        t = self.sd[name].float()
        return t.half().float() if self.q and t.dim() == 4 else t

    def conv_bn(self, x, prefix, stride=1, pad=0, relu=False, residual=None):
        y = F.conv2d(x, self.w(prefix + ".weight"), None, stride, pad)
        if self.q:  # the kernel folds FrozenBN into scale/shift (tolerance-checked against F.batch_norm)
            scale = self.sd[prefix + ".norm.weight"] * torch.rsqrt(self.sd[prefix + ".norm.running_var"] + 1e-5)
            shift = self.sd[prefix + ".norm.bias"] - self.sd[prefix + ".norm.running_mean"] * scale
            y = y * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
        else:
            y = F.batch_norm(y, self.sd[prefix + ".norm.running_mean"], self.sd[prefix + ".norm.running_var"],
                             self.sd[prefix + ".norm.weight"], self.sd[prefix + ".norm.bias"], False, 0.0, 1e-5)
        if residual is not None:
            y = y + residual
        if relu:
            y = F.relu(y)
        return self.rnd(y)

    def conv_bias(self, x, prefix, stride=1, pad=0, add=None, store=True):
        y = F.conv2d(x, self.w(prefix + ".weight"), self.sd[prefix + ".bias"].float(), stride, pad)
        if add is not None:
            y = y + add
        return self.rnd(y) if store else y


@torch.no_grad()
def forward_dense(sd: Dict[str, torch.Tensor], depth: int, batch: torch.Tensor, mode: str = "fp32",
                  num_threads: int | None = None) -> Dict[str, List[torch.Tensor]]:
    """batch: normalised padded N x 3 x H x W float32. Returns per-level features and head outputs (NCHW fp32)."""
    if num_threads:
        torch.set_num_threads(num_threads)
    net = _Net(sd, mode)
    x = net.rnd(batch)
    named = {}
    x = net.conv_bn(x, BU + "stem.conv1", 2, 3, relu=True)
    named["stem"] = x
    x = F.max_pool2d(x, 3, 2, 1)
    named["pool"] = x
    feats = {}
    for s, nblocks in zip(range(2, 6), STAGE_BLOCKS[depth]):
        for b in range(nblocks):
            pre = f"{BU}res{s}.{b}"
            stride = 2 if (b == 0 and s > 2) else 1
            sc = net.conv_bn(x, pre + ".shortcut", stride) if b == 0 else x
            y = net.conv_bn(x, pre + ".conv1", stride, relu=True)
            y = net.conv_bn(y, pre + ".conv2", 1, 1, relu=True)
            x = net.conv_bn(y, pre + ".conv3", relu=True, residual=sc)
            named[f"res{s}.{b}"] = x
        feats[s] = x
    # FPN top-down (detectron2 FPN.forward): prev = lateral(c5); p5 = output(prev); then lateral + nearest 2x
    prev = net.conv_bias(feats[5], "backbone.fpn_lateral5")
    P = {5: net.conv_bias(prev, "backbone.fpn_output5", 1, 1)}
    for i in (4, 3):
        up = F.interpolate(prev, scale_factor=2.0, mode="nearest")
        prev = net.conv_bias(feats[i], f"backbone.fpn_lateral{i}", add=up)
        P[i] = net.conv_bias(prev, f"backbone.fpn_output{i}", 1, 1)
    p6 = net.conv_bias(P[5], "backbone.top_block.p6", 2, 1)
    p7 = net.conv_bias(F.relu(p6), "backbone.top_block.p7", 2, 1)
    levels = [P[3], P[4], P[5], p6, p7]
    for i, f in enumerate(levels):
        named[f"p{3 + i}"] = f

    def tower(f, name):
        for i in range(4):
            raw = net.conv_bias(f, f"{HEAD}{name}.{3 * i}", 1, 1)  # stored fp16 in o16 mode, GN statistics on that
            f = F.relu(F.group_norm(raw, 32, sd[f"{HEAD}{name}.{3 * i + 1}.weight"].float(),
                                    sd[f"{HEAD}{name}.{3 * i + 1}.bias"].float(), 1e-5))
            f = net.rnd(f)
        return f

    out = {"features": levels, "logits": [], "reg": [], "ctr": [], "center": [], "named": named}
    for l, f in enumerate(levels):
        cls_t = tower(f, "cls_tower")
        ctr_t = tower(f, "center_tower")
        cor_t = tower(ctr_t, "corners_tower")  # CORNER_TOWER_ON_CENTER_TOWER
        named[f"cls_tower.l{l}"], named[f"center_tower.l{l}"], named[f"corners_tower.l{l}"] = cls_t, ctr_t, cor_t
        center = net.conv_bias(ctr_t, HEAD + "center_pred", 1, 1, store=False)
        delta = net.conv_bias(cor_t, HEAD + "corners_pred", 1, 1, store=False)
        scale = sd[f"{HEAD}scales.{l}.scale"].float()
        reg = (center.repeat(1, 4, 1, 1) + delta) * scale  # dafne.py:405-411
        out["logits"].append(net.conv_bias(cls_t, HEAD + "cls_logits", 1, 1, store=False))
        out["ctr"].append(net.conv_bias(cor_t, HEAD + "ctrness", 1, 1, store=False))  # CTR_ON_REG
        out["reg"].append(reg)
        out["center"].append(center * scale)
    return out
