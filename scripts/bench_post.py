"""Dev aid: time the post-processing (and the whole step) of one workload with a given build of the library.

  python scripts/bench_post.py [--lib dafne_b200/libdafne_b200_v1.so] [--workload r101_b32] [--steps 10]

Prints one JSON line: ms per post-processing call (CUDA events, head outputs of one forward resident), ms per full
step, the NMS work counters. Used for same-box A/B runs of kernel variants (make OUT=... OBJDIR=... EXTRA=-D...).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default="")
    ap.add_argument("--workload", default="r101_b32")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--batch", type=int, default=0)
    args = ap.parse_args()
    from dafne_b200 import _capi

    if args.lib:
        _capi.LIB_PATH = os.path.abspath(args.lib)
    import torch

    import bench
    from dafne_b200.engine import DafneEngine, DetectionWire

    cfg, spec, batch, H, W, what = bench.load_spec(args.workload)
    batch = args.batch or batch
    dev = torch.device("cuda:0")
    eng = DafneEngine(spec, dev)
    eng.load_state_dict(bench.synth_weights(args.workload, spec))
    g = torch.Generator().manual_seed(1234)
    sets = [torch.randint(0, 256, (batch, 3, H, W), dtype=torch.uint8, generator=g).to(dev) for _ in range(2)]
    sizes = [(H, W)] * batch
    cap = spec.post_nms_topk + 24
    wire = DetectionWire(batch, cap, dev)
    for i in range(3):
        eng.detect(sets[i % 2], sizes, None, True, cap, out=wire)
    torch.cuda.synchronize()

    def timed(fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    ms_post = timed(lambda i: eng.postprocess(sizes, None, True, cap, out=wire), args.steps)
    ms_step = timed(lambda i: eng.detect(sets[i % 2], sizes, None, True, cap, out=wire), args.steps)
    counts = wire.counts.cpu().tolist()
    pc = eng.post_counts()
    print(json.dumps({"lib": os.path.basename(_capi.LIB_PATH), "workload": args.workload, "batch": batch,
                      "env": {k: v for k, v in os.environ.items() if k.startswith("DAFNE_")},
                      "ms_post": ms_post, "ms_step": ms_step, "images_per_s": batch / ms_step * 1e3,
                      "nms": eng.nms_stats(), "nms_in": [c["nms_in"] for c in pc[:4]],
                      "nms_kept": [c["nms_kept"] for c in pc[:4]], "detections": counts[:4],
                      "checksum": float(wire.dets.double().sum().item())}))


if __name__ == "__main__":
    main()
