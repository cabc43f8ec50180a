"""Search (on the CPU, with the oracle) for the fixed synthetic input of the end-to-end IDENTITY gate
(tests/test_identity_gpu.py): an image seed for which no decision of the post-processing chain sits within the drift
of the fp16 dense path.

The reference computes the dense forward in fp32; this repository stores fp16 activations between layers, so head
outputs drift (measured: |d logit| ~ 1e-2, tests/test_dense_gpu.py) and every discontinuous decision -- score > 0.05,
membership of the per-level top-k, IoU > 0.1, the post-NMS cut -- can flip for a candidate that sits on the boundary.
SURVEY.md section 7 asks for inputs chosen, and documented, so that none does. A seed is accepted when
  (1) the fp32 oracle and the quantisation-matched oracle (o16) return the SAME detections (indices, classes, order), and
  (2) so do `trials` runs of the fp32 heads perturbed by uniform noise whose standard deviation is `noise` x the measured
      RMS fp32-vs-o16 drift (per tensor, per level) -- i.e. every decision has a margin of several times the drift.

Margins are MADE, not hoped for: for the fixed image the per-class bias of cls_logits (15 free numbers of the synthetic
recipe) is placed in the widest gap of that class's score distribution near the wanted detection count -- every
(location, class) has a bias value at which it crosses the 0.05 threshold; the bias goes to the middle of the largest
gap between consecutive crossing values. The seed search then only has to find an image whose NMS / ordering
decisions have margins too.

  python scripts/find_identity_input.py [--depth 50|101] [--twc 0|1] [--hw 256 320] [--per-class 4] [--tries 40]
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dafne_b200.spec import ModelSpec  # noqa: E402
from dafne_b200.weights import synthetic_state_dict  # noqa: E402
from oracle import model as omodel  # noqa: E402
from oracle import postprocess as opost  # noqa: E402


def post(spec, out, sizes, noise=None, rng=None):
    lg, rg, ct = ([t.numpy().copy() for t in out[k]] for k in ("logits", "reg", "ctr"))
    if noise is not None:
        for arrs, amp in ((lg, noise[0]), (rg, noise[1]), (ct, noise[2])):
            for l, a in enumerate(arrs):
                a += rng.uniform(-amp[l], amp[l], a.shape).astype(np.float32)
    return opost.postprocess(lg, rg, ct, spec.fpn_strides, sizes, None, score_thresh=spec.score_thresh,
                             pre_nms_topk=spec.pre_nms_topk, nms_thresh=spec.nms_thresh,
                             post_nms_topk=spec.post_nms_topk, sort_corners=spec.sort_corners,
                             thresh_with_ctr=spec.thresh_with_ctr, vehicle_merge=spec.vehicle_merge)


def calibrate_bias(spec, out, b0, per_class, window=4):
    """Per-class bias in the widest gap of the crossing values (see the module header). `out`: fp32 oracle heads computed
    with the uniform bias b0. Returns (biases [C], smallest half-gap in logit units)."""
    C_ = spec.num_classes
    thr_logit = float(np.log(spec.score_thresh / (1.0 - spec.score_thresh)))
    cross = [[] for _ in range(C_)]
    for lg, ct in zip(out["logits"], out["ctr"]):
        z = lg.numpy().astype(np.float64) - b0  # [N, C, H, W] without the bias
        if spec.thresh_with_ctr:
            # sqrt(cls * ctr) > t  <=>  cls > t^2 / ctr  <=>  z + b > logit(t^2 / ctr)
            c = 1.0 / (1.0 + np.exp(-ct.numpy().astype(np.float64)))
            need = spec.score_thresh ** 2 / c
            need = np.where(need < 1.0, np.log(np.clip(need, 1e-300, 1 - 1e-12) / (1.0 - np.clip(need, 0, 1 - 1e-12))), np.inf)
            bc = need - z  # broadcast over classes
        else:
            bc = thr_logit - z
        for k in range(C_):
            cross[k].append(bc[:, k].reshape(-1))
    biases, margin = [], np.inf
    for k in range(C_):
        v = np.sort(np.concatenate(cross[k]))  # ascending: the first `r` values pass when b lies above them
        lo, hi = max(per_class - window // 2, 1), per_class + window
        gaps = v[lo:hi + 1] - v[lo - 1:hi]
        j = int(np.argmax(gaps))
        biases.append(float(0.5 * (v[lo - 1 + j] + v[lo + j])))
        margin = min(margin, 0.5 * float(gaps[j]))
    return biases, margin


def same(a, b):
    """Identical detections: the same set of (candidate index, class) per image. (The ORDER of two detections whose
    scores differ by less than the drift may swap; that is checked separately by the test, where the gap allows.)"""
    for x, y in zip(a, b):
        if len(x["canon"]) != len(y["canon"]):
            return False
        ox, oy = np.argsort(x["canon"]), np.argsort(y["canon"])
        if not (np.array_equal(x["canon"][ox], y["canon"][oy]) and np.array_equal(x["pred_classes"][ox], y["pred_classes"][oy])):
            return False
    return True


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--depth", type=int, default=50)
    ap.add_argument("--twc", type=int, default=0)
    ap.add_argument("--classes", type=int, default=15)
    ap.add_argument("--hw", type=int, nargs=2, default=[256, 320])
    ap.add_argument("--bias", type=float, default=None)
    ap.add_argument("--per-class", type=int, default=0, help="calibrate a per-class bias for about this many candidates per class")
    ap.add_argument("--tries", type=int, default=40)
    ap.add_argument("--trials", type=int, default=12)
    ap.add_argument("--noise", type=float, default=3.0)
    ap.add_argument("--first-seed", type=int, default=1000)
    args = ap.parse_args()
    spec = ModelSpec(resnet_depth=args.depth, num_classes=args.classes, thresh_with_ctr=bool(args.twc))
    sd = synthetic_state_dict(spec, 0, cls_bias=args.bias)
    H, W = args.hw
    for seed in range(args.first_seed, args.first_seed + args.tries):
        g = torch.Generator().manual_seed(seed)
        img = torch.randint(0, 256, (3, H, W), dtype=torch.uint8, generator=g)
        batch, sizes = omodel.preprocess([img], spec.pixel_mean, spec.pixel_std)
        biases, margin = None, None
        if args.per_class:
            b0 = float(sd["proposal_generator.dafne_head.cls_logits.bias"][0])
            sd["proposal_generator.dafne_head.cls_logits.bias"] = torch.full((args.classes,), b0)
            biases, margin = calibrate_bias(spec, omodel.forward_dense(sd, args.depth, batch, "fp32"), b0, args.per_class)
            sd["proposal_generator.dafne_head.cls_logits.bias"] = torch.tensor(biases, dtype=torch.float32)
        o32 = omodel.forward_dense(sd, args.depth, batch, "fp32")
        o16 = omodel.forward_dense(sd, args.depth, batch, "o16")
        drift = [[float((a - b).abs().max()) for a, b in zip(o32[k], o16[k])] for k in ("logits", "reg", "ctr")]
        rms = [[float((a - b).pow(2).mean().sqrt()) for a, b in zip(o32[k], o16[k])] for k in ("logits", "reg", "ctr")]
        r32, r16 = post(spec, o32, sizes), post(spec, o16, sizes)
        n = len(r32[0]["scores"])
        ok = same(r32, r16) and n >= 10
        rng = np.random.default_rng(seed)
        t = 0
        while ok and t < args.trials:
            # uniform on [-a, a] has standard deviation a / sqrt(3)
            ok = same(r32, post(spec, o32, sizes, [[args.noise * d * 3 ** 0.5 for d in dr] for dr in rms], rng))
            t += 1
        print(f"seed {seed}: {n} detections, max drift logits {max(drift[0]):.4f} reg {max(drift[1]):.4f} "
              f"ctr {max(drift[2]):.4f} (rms {max(rms[0]):.4f} {max(rms[1]):.4f} {max(rms[2]):.4f}), o16 == fp32: {same(r32, r16)}, robust trials passed: {t}/{args.trials}",
              flush=True)
        if biases is not None:
            print(f"   threshold half-gap {margin:.4f} logit units; biases {[round(b, 5) for b in biases]}", flush=True)
        if ok:
            print(f"ACCEPTED seed={seed} depth={args.depth} twc={args.twc} hw={H}x{W} bias={args.bias} "
                  f"per_class_biases={[float(np.float32(b)) for b in biases] if biases else None}")
            return 0
    return 1


if __name__ == "__main__":
    sys.exit(main())
