"""Test-time augmentation throughput on one B200: the reference's DOTA-1.0 TTA recipe
(configs/pre-trained/dota-1.0_r101_ms.yaml:395-411 -- 9 scales x {plain, hflip, vflip} = 27 copies per image, batches of
3) through dafne_b200.tta.OneStageRCNNWithTTA, with the copies built on the host (PIL, like the reference's mapper) or
on the device (Pillow-exact resize kernel).

  python scripts/bench_tta.py [--depth 101] [--images 4] [--mapper device|host]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--depth", type=int, default=101)
    ap.add_argument("--images", type=int, default=4)
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--mapper", default="device", choices=["device", "host"])
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cross-image-batch", type=int, default=12, help="0 = the reference's per-image batches of 3")
    args = ap.parse_args()
    from dafne_b200 import tta
    from dafne_b200.config import get_cfg
    from dafne_b200.modeling import build_model

    cfg = get_cfg()
    cfg.merge_from_file(os.path.join(ROOT, "configs", "dota10_r101_ms.yaml" if args.depth == 101 else "dota10_r50_1024.yaml"))
    cfg.MODEL.DEVICE = "cuda:0"
    cfg.TEST.AUG.MIN_SIZES = [450, 500, 600, 700, 800, 900, 1000, 1100, 1200]
    cfg.TEST.AUG.MAX_SIZE = 1200
    model = build_model(cfg)
    mapper = tta.DotaDatasetMapperTTA(cfg, device="cuda:0" if args.mapper == "device" else None)
    wrapper = tta.OneStageRCNNWithTTA(cfg, model, tta_mapper=mapper, cross_image_batch=args.cross_image_batch)
    g = torch.Generator().manual_seed(0)
    imgs = [torch.randint(0, 256, (3, args.size, args.size), dtype=torch.uint8, generator=g) for _ in range(args.images)]
    inputs = [{"image": im, "height": args.size, "width": args.size} for im in imgs]
    # warm-up with the call shape that is timed: plans one engine per (scale, batch) and captures its graph -- the steady
    # state of a loop over a dataset, whose images repeat the same sizes
    out = wrapper(inputs)
    torch.cuda.synchronize()
    g2 = torch.Generator().manual_seed(1)
    inputs = [{"image": torch.randint(0, 256, (3, args.size, args.size), dtype=torch.uint8, generator=g2),
               "height": args.size, "width": args.size} for _ in range(args.images)]
    dts = []
    for _ in range(args.reps):
        t0 = time.perf_counter()
        outs = wrapper(inputs)
        torch.cuda.synchronize()
        dts.append(time.perf_counter() - t0)
    dt = min(dts)
    print(json.dumps({"op": "TTA, 27 copies per image (9 scales x 3 flips), union polygon NMS", "depth": args.depth,
                      "mapper": args.mapper, "images": args.images, "cross_image_batch": args.cross_image_batch,
                      "s_per_call": [round(x, 4) for x in dts], "images_per_s": args.images / dt,
                      "copies_per_s": 27 * args.images / dt, "ms_per_image": dt / args.images * 1e3,
                      "detections": [len(o["instances"]) for o in outs],
                      "warmup_detections": len(out[0]["instances"])}))


if __name__ == "__main__":
    main()
