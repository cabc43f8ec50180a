"""Patch-merge polygon NMS: device (dafne_b200.merge) vs the reference's CPU algorithm restated in
oracle/merge_nms.py (numpy + the double-precision C polygon IoU), on DOTA-like per-class detection lists.

  python scripts/bench_merge_nms.py [--images 32] [--boxes 1500]

Prints one JSON line: boxes/s through the device call (host buffers in and out, allocation and copies included) and
through the CPU restatement, the number of polygon IoUs the reference would evaluate, and that the results agree.
Measurement tool only: the oracle import is the checker / CPU baseline, as in bench.py's cpu_baseline leg.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rects(rng, n, extent, wmin, wmax, aspect):
    cx, cy = rng.uniform(0, extent, n), rng.uniform(0, extent, n)
    w = rng.uniform(wmin, wmax, n)
    h = w / aspect
    a = rng.uniform(0, np.pi, n)
    dx = np.stack([-w, w, w, -w], 1) / 2
    dy = np.stack([-h, -h, h, h], 1) / 2
    x = cx[:, None] + dx * np.cos(a)[:, None] - dy * np.sin(a)[:, None]
    y = cy[:, None] + dx * np.sin(a)[:, None] + dy * np.cos(a)[:, None]
    return np.stack([x, y], 2).reshape(n, 8)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=32)
    ap.add_argument("--boxes", type=int, default=1500)
    ap.add_argument("--cpu-images", type=int, default=4)
    args = ap.parse_args()
    from dafne_b200 import merge
    from oracle import merge_nms as omerge

    rng = np.random.default_rng(0)
    lists = []
    for _ in range(args.images):
        n = args.boxes
        b = np.round(rects(rng, n, 3000, 20, 140, 4.0), 1)
        b[n // 2:] = b[: n - n // 2] + rng.normal(0, 2.0, (n - n // 2, 8))  # every object seen by two patches
        s = rng.permutation(n) / n * 0.9 + 0.05 + rng.uniform(0, 1e-5, n)
        lists.append(np.concatenate([b, s[:, None]], 1))
    merge.nms_many(lists[:2], 0.1)  # warm-up (context, module load)
    t0 = time.perf_counter()
    got = merge.nms_many(lists, 0.1)
    t_dev = time.perf_counter() - t0
    t0 = time.perf_counter()
    want = [omerge.py_cpu_nms_poly_fast(d, 0.1) for d in lists[: args.cpu_images]]
    t_cpu = time.perf_counter() - t0
    print(json.dumps({
        "op": "patch-merge polygon NMS, float64 (py_cpu_nms_poly_fast)", "images": args.images, "boxes_per_image": args.boxes,
        "device_boxes_per_s": args.images * args.boxes / t_dev, "device_ms": t_dev * 1e3,
        "cpu_port_boxes_per_s": args.cpu_images * args.boxes / t_cpu, "cpu_images": args.cpu_images,
        "speedup": (args.images * args.boxes / t_dev) / (args.cpu_images * args.boxes / t_cpu),
        "kept_per_image": [len(k) for k in got[:4]], "agree": got[: args.cpu_images] == want}))


if __name__ == "__main__":
    main()
