"""Small-batch latency of one step (dense forward + post-processing), eager launches vs the captured CUDA graph.

  python scripts/bench_latency.py [--depth 101] [--batches 1 3 8] [--size 1024] [--iters 30]

Prints one JSON line per batch size: ms per step issued eagerly (about 200 launches) and replayed as one graph
(dafne_graph_capture / dafne_graph_launch), both timed with CUDA events over back-to-back steps, and the host time the
issuing thread spends per step.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--depth", type=int, default=101)
    ap.add_argument("--batches", type=int, nargs="+", default=[1, 3, 8])
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=30)
    args = ap.parse_args()
    import torch

    import bench
    from dafne_b200.engine import DafneEngine, DetectionWire

    wl = "r101_b32" if args.depth == 101 else "r50_b8"
    cfg, spec, _, _, _, _ = bench.load_spec(wl)
    dev = torch.device("cuda:0")
    sd = bench.synth_weights(wl, spec)
    for b in args.batches:
        eng = DafneEngine(spec, dev)
        eng.load_state_dict(sd)
        g = torch.Generator().manual_seed(b)
        img = torch.randint(0, 256, (b, 3, args.size, args.size), dtype=torch.uint8, generator=g).to(dev)
        sizes = [(args.size, args.size)] * b
        cap = spec.post_nms_topk + 24
        wire = DetectionWire(b, cap, dev)
        for _ in range(3):
            eng.detect(img, sizes, None, True, cap, out=wire)
        torch.cuda.synchronize()

        def timed(fn):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e0.record()
            for _ in range(args.iters):
                fn()
            e1.record()
            host = (time.perf_counter() - t0) / args.iters * 1e3
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / args.iters, host

        eager_ms, eager_host = timed(lambda: eng.detect(img, sizes, None, True, cap, out=wire))
        ref = wire.dets.clone()
        eng.capture(img, sizes, None, True, cap, out=wire)
        graph_ms, graph_host = timed(eng.replay)
        same = bool(torch.equal(ref, wire.dets))
        print(json.dumps({"depth": args.depth, "batch": b, "size": args.size, "eager_ms_per_step": eager_ms,
                          "graph_ms_per_step": graph_ms, "eager_host_ms_per_step": eager_host,
                          "graph_host_ms_per_step": graph_host, "images_per_s_eager": b / eager_ms * 1e3,
                          "images_per_s_graph": b / graph_ms * 1e3, "identical_results": same}), flush=True)
        eng.close()


if __name__ == "__main__":
    main()
