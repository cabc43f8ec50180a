"""Comparison baseline on the same B200: the reference's STRUCTURE in eager PyTorch (cuDNN / ATen library kernels).

detectron2 and poly_nms cannot be installed offline, so this is not the reference itself; it is the same module graph
built from torch.nn layers the way detectron2 v0.5 builds it (nn.Conv2d, FrozenBN as a per-channel affine, nn.GroupNorm,
max_pool2d, nearest interpolate) followed by the reference's post-processing structure: a Python loop over levels and
images with sigmoid / threshold / nonzero / topk (dafne/modeling/dafne/dafne_outputs.py:792-905) and a per-image NMS with
the reference's host round trip (dafne/modeling/nms/nms.py:86-91 -- boxes.cpu().numpy() -> poly_gpu_nms(dets, thr, dev)),
where poly_gpu_nms is this repository's drop-in `dafne_poly_nms_host` (INTEGRATION.md seam 1).

  python scripts/bench_torch_eager.py [--depth 50] [--batch 8] [--steps 10] [--warmup 3] [--modes fp32,tf32,fp16cl]

Prints one JSON line per mode: images/s of the dense forward alone and of forward + post-processing.
Self-contained on purpose: it does not import oracle/ (test infrastructure) and it is not part of the product path.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

BU = "backbone.bottom_up."
HEAD = "proposal_generator.dafne_head."
STAGE_BLOCKS = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3)}


class FrozenBN(nn.Module):
    def __init__(self, sd, prefix):
        super().__init__()
        scale = sd[prefix + ".weight"] * torch.rsqrt(sd[prefix + ".running_var"] + 1e-5)
        self.register_buffer("scale", scale.view(1, -1, 1, 1).clone())
        self.register_buffer("shift", (sd[prefix + ".bias"] - sd[prefix + ".running_mean"] * scale).view(1, -1, 1, 1).clone())

    def forward(self, x):
        return x * self.scale + self.shift


def conv(sd, prefix, stride=1, pad=0, bias=True):
    w = sd[prefix + ".weight"]
    m = nn.Conv2d(w.shape[1], w.shape[0], w.shape[2], stride, pad, bias=bias)
    m.weight.data.copy_(w)
    if bias:
        m.bias.data.copy_(sd[prefix + ".bias"])
    return m


class Bottleneck(nn.Module):
    def __init__(self, sd, pre, stride, has_sc):
        super().__init__()
        self.sc = nn.Sequential(conv(sd, pre + ".shortcut", stride, 0, False), FrozenBN(sd, pre + ".shortcut.norm")) if has_sc else None
        self.c1, self.b1 = conv(sd, pre + ".conv1", stride, 0, False), FrozenBN(sd, pre + ".conv1.norm")
        self.c2, self.b2 = conv(sd, pre + ".conv2", 1, 1, False), FrozenBN(sd, pre + ".conv2.norm")
        self.c3, self.b3 = conv(sd, pre + ".conv3", 1, 0, False), FrozenBN(sd, pre + ".conv3.norm")

    def forward(self, x):
        s = self.sc(x) if self.sc is not None else x
        y = F.relu_(self.b1(self.c1(x)))
        y = F.relu_(self.b2(self.c2(y)))
        return F.relu_(self.b3(self.c3(y)) + s)


class EagerDafne(nn.Module):
    def __init__(self, sd, depth, num_classes):
        super().__init__()
        self.stem, self.stem_bn = conv(sd, BU + "stem.conv1", 2, 3, False), FrozenBN(sd, BU + "stem.conv1.norm")
        self.stages = nn.ModuleList()
        for s, nb in zip(range(2, 6), STAGE_BLOCKS[depth]):
            self.stages.append(nn.Sequential(*[Bottleneck(sd, f"{BU}res{s}.{b}", 2 if (b == 0 and s > 2) else 1, b == 0)
                                               for b in range(nb)]))
        self.lat = nn.ModuleList([conv(sd, f"backbone.fpn_lateral{i}") for i in (3, 4, 5)])
        self.out = nn.ModuleList([conv(sd, f"backbone.fpn_output{i}", 1, 1) for i in (3, 4, 5)])
        self.p6, self.p7 = conv(sd, "backbone.top_block.p6", 2, 1), conv(sd, "backbone.top_block.p7", 2, 1)

        def tower(name):
            layers = []
            for i in range(4):
                layers.append(conv(sd, f"{HEAD}{name}.{3 * i}", 1, 1))
                gn = nn.GroupNorm(32, 256)
                gn.weight.data.copy_(sd[f"{HEAD}{name}.{3 * i + 1}.weight"])
                gn.bias.data.copy_(sd[f"{HEAD}{name}.{3 * i + 1}.bias"])
                layers += [gn, nn.ReLU()]
            return nn.Sequential(*layers)

        self.cls_tower, self.center_tower, self.corners_tower = tower("cls_tower"), tower("center_tower"), tower("corners_tower")
        self.cls_logits, self.ctrness = conv(sd, HEAD + "cls_logits", 1, 1), conv(sd, HEAD + "ctrness", 1, 1)
        self.corners_pred, self.center_pred = conv(sd, HEAD + "corners_pred", 1, 1), conv(sd, HEAD + "center_pred", 1, 1)
        self.scales = [float(sd[f"{HEAD}scales.{l}.scale"]) for l in range(5)]

    def forward(self, x):
        x = F.max_pool2d(F.relu_(self.stem_bn(self.stem(x))), 3, 2, 1)
        feats = []
        for i, st in enumerate(self.stages):
            x = st(x)
            if i >= 1:
                feats.append(x)
        prev = self.lat[2](feats[2])
        P = {5: self.out[2](prev)}
        for i in (1, 0):
            prev = self.lat[i](feats[i]) + F.interpolate(prev, scale_factor=2.0, mode="nearest")
            P[3 + i] = self.out[i](prev)
        p6 = self.p6(P[5])
        p7 = self.p7(F.relu(p6))
        logits, reg, ctr = [], [], []
        for l, f in enumerate([P[3], P[4], P[5], p6, p7]):
            cls_t = self.cls_tower(f)
            ctr_t = self.center_tower(f)
            cor_t = self.corners_tower(ctr_t)
            center = self.center_pred(ctr_t)
            delta = self.corners_pred(cor_t)
            reg.append((center.repeat(1, 4, 1, 1) + delta) * self.scales[l])
            logits.append(self.cls_logits(cls_t))
            ctr.append(self.ctrness(cor_t))
        return logits, reg, ctr


def postprocess_reference_structure(logits, reg, ctr, strides, nms_host, thr=0.05, topk=2000, nms_thr=0.1, post_topk=1000):
    """Python loops + ATen calls + per-image host round trip, like dafne_outputs.py:733-925 / nms.py:37-92."""
    N = logits[0].shape[0]
    per_image = [[] for _ in range(N)]
    for l, (lg, rg, ct) in enumerate(zip(logits, reg, ctr)):
        _, C, H, W = lg.shape
        s = strides[l]
        ys, xs = torch.meshgrid(torch.arange(H, device=lg.device), torch.arange(W, device=lg.device), indexing="ij")
        loc = torch.stack([xs.reshape(-1) * s + s // 2, ys.reshape(-1) * s + s // 2], 1).float()
        cls = lg.float().permute(0, 2, 3, 1).reshape(N, -1, C).sigmoid()
        rc = (rg.float() * s).permute(0, 2, 3, 1).reshape(N, -1, 8)
        cn = ct.float().permute(0, 2, 3, 1).reshape(N, -1).sigmoid()
        cand = cls > thr
        k = cand.reshape(N, -1).sum(1).clamp(max=topk)
        score = (cls * cn[:, :, None]).sqrt()
        for i in range(N):
            idx = cand[i].nonzero()
            sc = score[i][cand[i]]
            if sc.numel() > k[i].item():  # host sync, like dafne_outputs.py:851
                sc, top = sc.topk(int(k[i]), sorted=False)
                idx = idx[top]
            poly = loc[idx[:, 0]].repeat(1, 4) + rc[i][idx[:, 0]]
            per_image[i].append((poly, sc, idx[:, 1]))
    n_det = 0
    for i in range(N):
        poly = torch.cat([p[0] for p in per_image[i]])
        sc = torch.cat([p[1] for p in per_image[i]])
        cl = torch.cat([p[2] for p in per_image[i]]).clone()
        if poly.numel() == 0:
            continue
        cl[cl == 5] = 4
        off = cl.to(poly) * (poly.max() - poly.min() + 1)
        dets = torch.cat([poly + off[:, None], sc[:, None]], 1).cpu().numpy()  # the reference's D2H round trip
        keep = nms_host(dets, nms_thr, 0)
        keep = keep[:post_topk]
        n_det += len(keep)
    return n_det


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--depth", type=int, default=50)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--modes", default="fp32,tf32,fp16cl")
    ap.add_argument("--size", type=int, default=1024)
    args = ap.parse_args()

    from dafne_b200.modeling import poly_gpu_nms
    from dafne_b200.spec import ModelSpec
    from dafne_b200.weights import synthetic_state_dict

    spec = ModelSpec(resnet_depth=args.depth, num_classes=15)
    sd = synthetic_state_dict(spec, 0)
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(1234)
    imgs = torch.randint(0, 256, (args.batch, 3, args.size, args.size), dtype=torch.uint8, generator=g).to(dev)
    mean = torch.tensor(spec.pixel_mean, device=dev).view(1, 3, 1, 1)
    std = torch.tensor(spec.pixel_std, device=dev).view(1, 3, 1, 1)

    def nms_host(dets, thr, dev_id):
        return poly_gpu_nms(np.ascontiguousarray(dets, dtype=np.float32), thr, dev_id)

    for mode in args.modes.split(","):
        torch.backends.cudnn.allow_tf32 = mode != "fp32"
        torch.backends.cuda.matmul.allow_tf32 = mode != "fp32"
        torch.backends.cudnn.benchmark = True
        model = EagerDafne(sd, args.depth, 15).to(dev).eval()
        if mode == "fp16cl":
            model = model.half().to(memory_format=torch.channels_last)

        @torch.no_grad()
        def fwd():
            x = (imgs.float() - mean) / std
            if mode == "fp16cl":
                x = x.half().contiguous(memory_format=torch.channels_last)
            return model(x)

        def timed(fn, steps):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / steps

        n_det = [0]

        def full():
            lg, rg, ct = fwd()
            n_det[0] = postprocess_reference_structure(lg, rg, ct, spec.fpn_strides, nms_host)

        for _ in range(args.warmup):
            fwd()
        ms_fwd = timed(fwd, args.steps)
        full()
        ms_full = timed(full, max(2, args.steps // 3))
        print(json.dumps({
            "baseline": "torch-eager (cuDNN/ATen) graph + reference-structured post-processing", "mode": mode,
            "depth": args.depth, "batch": args.batch, "size": args.size,
            "forward_ms": ms_fwd, "forward_images_per_s": args.batch / ms_fwd * 1e3,
            "full_ms": ms_full, "full_images_per_s": args.batch / ms_full * 1e3, "detections": n_det[0],
            "torch": torch.__version__, "cudnn": torch.backends.cudnn.version()}), flush=True)
        del model
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
