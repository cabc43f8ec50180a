"""Dev aid: sustained TFLOP/s of a compute-bound 1x1 convolution (default M = 65 536, K = N = 4096) on the
single-CTA kernel (conv_tc.cu) and on the CTA-pair kernel (pair_tc.cu, tcgen05.mma.cta_group::2), each launched back to
back for about a second so that the board's power cap, not the burst clock, sets the number.

  python scripts/bench_pair_vs_single.py [--images 16] [--hw 64] [--k 4096] [--n 4096] [--reps 400]
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=16)
    ap.add_argument("--hw", type=int, default=64)
    ap.add_argument("--k", type=int, default=4096)
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=400)
    a = ap.parse_args()
    a.m = a.images * a.hw * a.hw
    from dafne_b200 import _capi

    lib = _capi.lib()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    x = torch.randn(a.m, a.k, generator=g).half().to(dev)
    w = (torch.randn(a.n, a.k, generator=g) / a.k ** 0.5).half().to(dev)
    sc = torch.ones(a.n, device=dev)
    sh = torch.zeros(a.n, device=dev)
    out = torch.empty(a.m, a.n, device=dev, dtype=torch.float16)
    vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    flops = 2.0 * a.m * a.k * a.n

    def single():
        _capi.check(lib.dafne_conv_nhwc(vp(x), a.images, a.hw, a.hw, a.k, vp(w), a.n, 1, 1, vp(sc), vp(sh), 1, None, 0, 0, 0, None,
                                        vp(out), None, 0, st), "dafne_conv_nhwc")

    def pair():
        _capi.check(lib.dafne_conv1x1_pair_nhwc(vp(x), a.m, a.k, vp(w), a.n, vp(sc), vp(sh), 1, None, vp(out), st),
                    "dafne_conv1x1_pair_nhwc")

    res = {}
    for name, fn in (("single_cta", single), ("cta_pair", pair), ("single_cta_again", single), ("cta_pair_again", pair)):
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        res[name] = {"ms": round(ms, 4), "tflops": round(flops / ms / 1e9, 1)}
    print(json.dumps({"m": a.m, "k": a.k, "n": a.n, "reps": a.reps, **res}))


if __name__ == "__main__":
    main()
