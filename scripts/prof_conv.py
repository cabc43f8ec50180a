"""Profiling target: the dominant kernel alone (head-tower 3x3 256->256 conv + bias + GN statistics at the p3 size of
the bench workload), a few launches through the C ABI. Run under `ncu --set full -k regex:conv_tc` on the GPU box.

  python scripts/prof_conv.py [N H W Cin Cout k stride reps residual gn_sums]
"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dafne_b200 import _capi  # noqa: E402


def main():
    a = [int(v) for v in sys.argv[1:]]
    N, H, W, Cin, Cout, k, stride, reps, use_res, use_gn = (a + [8, 128, 128, 256, 256, 3, 1, 5, 0, 1][len(a):])[:10]
    lib = _capi.lib()
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(0)
    x = torch.randn(N, H, W, Cin, generator=g).half().to(dev)
    w = (torch.randn(Cout, k, k, Cin, generator=g) / (Cin * k * k) ** 0.5).half().to(dev)
    shift = torch.randn(Cout, generator=g).to(dev)
    pad = k // 2
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    out = torch.empty(N, Ho, Wo, Cout, dtype=torch.float16, device=dev)
    sums = torch.zeros(N, Cout // 8, 2, dtype=torch.int64, device=dev)
    scale = (torch.rand(Cout, generator=g) + 0.5).to(dev)
    res = torch.randn(N, Ho, Wo, Cout, generator=g).half().to(dev) if use_res else None
    flush = torch.empty(256 * 2**20, dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    times = []
    for _ in range(reps):
        flush.zero_()  # evict the 126 MB L2 between launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        st = lib.dafne_conv_nhwc(x.data_ptr(), N, H, W, Cin, w.data_ptr(), Cout, k, stride, None if use_gn else scale.data_ptr(),
                                 shift.data_ptr(), 1 if use_res else 0, res.data_ptr() if use_res else None,
                                 Ho if use_res else 0, Wo if use_res else 0, 0, sums.data_ptr() if use_gn else None,
                                 out.data_ptr(), None, 0, s)
        e1.record()
        assert st == 0, _capi.last_error()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    fl = 2.0 * N * Ho * Wo * Cout * k * k * Cin
    best = min(times[1:]) if len(times) > 1 else times[0]
    by = (N * H * W * Cin + Cout * k * k * Cin + N * Ho * Wo * Cout * (2 if use_res else 1)) * 2
    print(f"conv N={N} {H}x{W} {Cin}->{Cout} k{k} s{stride} res={use_res} gn={use_gn}: best {best * 1e3:.1f} us, "
          f"{fl / best / 1e9:.1f} TFLOP/s, {by / best / 1e6:.0f} GB/s "
          f"(all: {[round(t * 1e3, 1) for t in times]})")


if __name__ == "__main__":
    main()
