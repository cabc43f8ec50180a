"""Where a TTA call spends its time (one B200): mapper (27 copies per image), the batched forwards, mapping the corners
back, the union polygon NMS. Synchronises between phases, so the sum exceeds the pipelined call of bench_tta.py.

  python scripts/profile_tta_phases.py [--images 4]
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
    ap.add_argument("--images", type=int, default=4)
    ap.add_argument("--size", type=int, default=1024)
    args = ap.parse_args()
    from dafne_b200 import tta
    from dafne_b200.config import get_cfg
    from dafne_b200.modeling import build_model

    cfg = get_cfg()
    cfg.merge_from_file(os.path.join(ROOT, "configs", "dota10_r101_ms.yaml"))
    cfg.MODEL.DEVICE = "cuda:0"
    cfg.TEST.AUG.MIN_SIZES = [450, 500, 600, 700, 800, 900, 1000, 1100, 1200]
    cfg.TEST.AUG.MAX_SIZE = 1200
    model = build_model(cfg)
    mapper = tta.DotaDatasetMapperTTA(cfg, device="cuda:0")
    w = tta.OneStageRCNNWithTTA(cfg, model, tta_mapper=mapper)
    g = torch.Generator().manual_seed(0)
    inputs = [{"image": torch.randint(0, 256, (3, args.size, args.size), dtype=torch.uint8, generator=g),
               "height": args.size, "width": args.size} for _ in range(args.images)]
    w(inputs)
    torch.cuda.synchronize()

    def timed(fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        return r, (time.perf_counter() - t0) * 1e3

    per_image, t_map = timed(lambda: [w._get_augmented_inputs(x) for x in inputs])
    groups = {}
    for k, (aug, _) in enumerate(per_image):
        for j, a in enumerate(aug):
            groups.setdefault(tuple(a["image"].shape), []).append((k, j, a))
    model.use_cuda_graphs = True

    def fwd():
        launched, owners = [], []
        for items in groups.values():
            for s0 in range(0, len(items), w.cross_image_batch):
                chunk = items[s0:s0 + w.cross_image_batch]
                launched.append(model._launch([a for _, _, a in chunk], do_postprocess=False))
                owners.append([(k, j) for k, j, _ in chunk])
        return launched, owners

    (launched, owners), t_launch = timed(fwd)
    res, t_collect = timed(lambda: model._collect(launched))
    model.use_cuda_graphs = False
    outputs = [[None] * len(aug) for aug, _ in per_image]
    for r, own in zip(res, owners):
        for o, (k, j) in zip(r, own):
            outputs[k][j] = o
    inst, t_back = timed(lambda: [w._to_original_frame(outputs[k], per_image[k][1]) for k in range(len(inputs))])
    merged, t_merge = timed(lambda: [w._merge_detections(i) for i in inst])
    n = args.images
    print(json.dumps({"images": n, "ms_per_image": {"mapper": t_map / n, "forwards_enqueue+run": t_launch / n,
                                                    "collect": t_collect / n, "corners_back": t_back / n,
                                                    "union_nms": t_merge / n},
                      "boxes_into_union_nms": [len(i) for i in inst], "kept": [len(m) for m in merged]}))


if __name__ == "__main__":
    main()
