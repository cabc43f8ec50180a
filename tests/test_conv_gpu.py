"""GPU parity of the tcgen05 implicit-GEMM convolution against a plain PyTorch fp32 reference of the same op.

The kernel computes fp16 x fp16 -> fp32 accumulate and stores fp16 (or fp32 for the small prediction convs), so the
reference is torch conv2d in fp32 (TF32 off) on the SAME fp16-rounded inputs and weights; the tolerance is one fp16
rounding of the output plus accumulation-order noise: |err| <= 2e-3 * max|ref| + 2e-3 * |ref|.
"""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

CASES = [
    # name, N, H, W, Cin, Cout, k, stride, opts
    ("1x1_64_64", 2, 32, 32, 64, 64, 1, 1, {}),
    ("3x3_256_256_bias_gn", 1, 64, 64, 256, 256, 3, 1, {"shift": True, "gn": True}),
    ("1x1s2_256_128_bn_relu", 2, 64, 64, 256, 128, 1, 2, {"scale": True, "shift": True, "relu": True}),
    ("3x3s2_256_256_p6", 2, 32, 32, 256, 256, 3, 2, {"shift": True}),
    ("3x3s2_odd25", 1, 25, 25, 256, 256, 3, 2, {"shift": True}),
    ("1x1_64_256_residual_relu", 2, 32, 32, 64, 256, 1, 1, {"scale": True, "shift": True, "relu": True, "res": 0}),
    ("1x1_512_256_fpn_upsample_add", 1, 32, 32, 512, 256, 1, 1, {"shift": True, "res": 1}),
    ("3x3_256_15_pred_f32", 2, 32, 32, 256, 15, 3, 1, {"shift": True, "f32": 16}),
    ("3x3_256_9_pred_f32", 1, 16, 16, 256, 9, 3, 1, {"shift": True, "f32": 16}),
    ("3x3_odd50_gn", 1, 50, 50, 256, 256, 3, 1, {"shift": True, "gn": True}),
    ("3x3_8x8_n3_gn", 3, 8, 8, 256, 256, 3, 1, {"shift": True, "gn": True}),
    ("3x3_4x4_n5_gn", 5, 4, 4, 256, 256, 3, 1, {"shift": True, "gn": True}),
    ("3x3_7x7_n2", 2, 7, 7, 256, 256, 3, 1, {"shift": True}),
    ("3x3_128_128", 1, 64, 64, 128, 128, 3, 1, {"scale": True, "shift": True, "relu": True}),
    ("3x3_512_512", 1, 32, 32, 512, 512, 3, 1, {"scale": True, "shift": True, "relu": True}),
    ("1x1_1024_2048s2", 1, 16, 16, 1024, 2048, 1, 2, {"scale": True, "shift": True}),
    ("3x3_256_256_multiwave", 2, 128, 128, 256, 256, 3, 1, {"shift": True, "gn": True}),
    # residual through the TMA ring (res_shift 0): ring wrap-around over many tiles per CTA, several n tiles,
    # both epilogue configurations (K <= 256: two warpgroups; deeper K: one), every tile width, ragged edges
    ("1x1_128_512_res_multiwave", 4, 128, 128, 128, 512, 1, 1, {"scale": True, "shift": True, "relu": True, "res": 0}),
    ("1x1_512_2048_res_wgs1", 2, 64, 64, 512, 2048, 1, 1, {"scale": True, "shift": True, "relu": True, "res": 0}),
    ("1x1_256_128_res_odd50", 2, 50, 50, 256, 128, 1, 1, {"scale": True, "shift": True, "relu": True, "res": 0}),
    ("3x3_128_64_res_odd40", 1, 40, 40, 128, 64, 3, 1, {"scale": True, "shift": True, "res": 0}),
    ("1x1_64_64_res_n3_30", 3, 30, 30, 64, 64, 1, 1, {"shift": True, "relu": True, "res": 0}),
    ("1x1_256_1024_res_multiwave", 8, 64, 64, 256, 1024, 1, 1, {"scale": True, "shift": True, "relu": True, "res": 0}),
    ("1x1_64_256_shortcut_multiwave", 2, 256, 256, 64, 256, 1, 1, {"scale": True, "shift": True}),
    # row-shared taps (3x3 stride 1, N tile <= 64): 8 x 16 pixel tiles, ragged sizes, several images, many tiles per CTA
    ("3x3_64_64_rowshared_multiwave", 2, 256, 256, 64, 64, 3, 1, {"scale": True, "shift": True, "relu": True}),
    ("3x3_256_15_pred_odd_n3", 3, 25, 42, 256, 15, 3, 1, {"shift": True, "f32": 16}),
    ("3x3_256_2_pred_7x11", 2, 7, 11, 256, 2, 3, 1, {"shift": True, "f32": 16}),
    ("3x3_256_32_pred_f32_ld32", 1, 40, 24, 256, 32, 3, 1, {"shift": True, "f32": 32}),
    ("3x3_64_64_res_rowshared", 1, 50, 70, 64, 64, 3, 1, {"scale": True, "shift": True, "relu": True, "res": 0}),
]


def run_conv_case(lib, name, N, H, W, Cin, Cout, k, stride, opts, seed=0):
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = (torch.randn(N, Cin, H, W, generator=g) * 1.0).half()
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).half()
    scale = (torch.rand(Cout, generator=g) + 0.5) if opts.get("scale") else None
    shift = torch.randn(Cout, generator=g) if opts.get("shift") else None
    pad = k // 2
    Ho = (H + 2 * pad - k) // stride + 1
    Wo = (W + 2 * pad - k) // stride + 1
    res = None
    rs = opts.get("res")
    if rs is not None:
        rH, rW = (Ho + (1 << rs) - 1) >> rs, (Wo + (1 << rs) - 1) >> rs
        res = torch.randn(N, Cout, rH, rW, generator=g).half()

    x_d = x.to(dev).permute(0, 2, 3, 1).contiguous()
    w_d = w.to(dev).permute(0, 2, 3, 1).contiguous()  # [Cout][kh][kw][Cin]
    scale_d = scale.to(dev) if scale is not None else None
    shift_d = shift.to(dev) if shift is not None else None
    res_d = res.to(dev).permute(0, 2, 3, 1).contiguous() if res is not None else None
    f32_ld = opts.get("f32")
    if f32_ld:
        out_d = torch.full((N, Ho, Wo, f32_ld), float("nan"), device=dev, dtype=torch.float32)
    else:
        out_d = torch.full((N, Ho, Wo, Cout), float("nan"), device=dev, dtype=torch.float16)
    sums_d = torch.zeros(N, Cout // 8, 2, device=dev, dtype=torch.int64) if opts.get("gn") else None

    def p(t):
        return C.c_void_p(0 if t is None else t.data_ptr())

    st = lib.dafne_conv_nhwc(
        p(x_d), N, H, W, Cin, p(w_d), Cout, k, stride, p(scale_d), p(shift_d), int(bool(opts.get("relu"))),
        p(res_d), 0 if res is None else res.shape[2], 0 if res is None else res.shape[3], rs or 0, p(sums_d),
        p(None if f32_ld else out_d), p(out_d if f32_ld else None), f32_ld or 0,
        C.c_void_p(torch.cuda.current_stream().cuda_stream),
    )
    assert st == 0, lib.dafne_last_error()
    torch.cuda.synchronize()

    # plain PyTorch fp32 reference on the same fp16-rounded operands
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref = F.conv2d(x.to(dev).float(), w.to(dev).float(), None, stride, pad)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    if scale is not None:
        ref = ref * scale.to(dev).view(1, -1, 1, 1)
    if shift is not None:
        ref = ref + shift.to(dev).view(1, -1, 1, 1)
    if res is not None:
        r = res.to(dev).float()
        if rs:
            r = F.interpolate(r, scale_factor=2, mode="nearest")[:, :, :Ho, :Wo]
        ref = ref + r
    if opts.get("relu"):
        ref = ref.relu()
    ref = ref.permute(0, 2, 3, 1)
    got = out_d.float()[..., :Cout]
    assert torch.isfinite(got).all(), f"{name}: non-finite / unwritten outputs"
    err = (got - ref).abs()
    tol = 2e-3 * ref.abs().max() + 2e-3 * ref.abs()
    bad = (err > tol).sum().item()
    assert bad == 0, f"{name}: {bad} of {err.numel()} outside tolerance, max err {err.max().item():.4g}"
    if f32_ld and f32_ld > Cout:
        assert (out_d[..., Cout:] == 0).all(), f"{name}: padded output channels must be exactly 0"
    if sums_d is not None:
        q = out_d.float().reshape(N, Ho * Wo, Cout // 8, 8)
        s1 = q.sum(dim=(1, 3))
        s2 = (q * q).sum(dim=(1, 3))
        # statistics are 64-bit fixed point (sum * 2^20, sumsq * 2^12): order-independent, hence reproducible
        got1 = sums_d[..., 0].double() / 2.0 ** 20
        got2 = sums_d[..., 1].double() / 2.0 ** 12
        assert torch.allclose(got1, s1.double(), rtol=1e-5, atol=1e-2), f"{name}: GN sum mismatch"
        assert torch.allclose(got2, s2.double(), rtol=1e-5, atol=1e-2), f"{name}: GN sumsq mismatch"
    return err.max().item()


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_conv_tc_matches_torch_fp32(case):
    from dafne_b200 import _capi

    run_conv_case(_capi.lib(), *case)


@pytest.mark.gpu
def test_gn_relu_matches_torch():
    from dafne_b200 import _capi

    lib = _capi.lib()
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(1)
    N, H, W, Cc = 3, 20, 12, 256
    x = (torch.randn(N, Cc, H, W, generator=g) * 2 + 0.5).half().to(dev)
    gamma = (torch.rand(Cc, generator=g) + 0.5).to(dev)
    beta = torch.randn(Cc, generator=g).to(dev)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    q = x_nhwc.float().reshape(N, H * W, 32, 8)
    sums = torch.stack([(q.double().sum(dim=(1, 3)) * 2.0 ** 20).round(), ((q.double() ** 2).sum(dim=(1, 3)) * 2.0 ** 12).round()],
                       dim=-1).to(torch.int64).contiguous()  # the fixed-point form dafne_conv_nhwc accumulates
    out = torch.empty_like(x_nhwc)
    st = lib.dafne_gn_relu_nhwc(x_nhwc.data_ptr(), out.data_ptr(), N, H * W, Cc, 32, sums.data_ptr(),
                                gamma.data_ptr(), beta.data_ptr(), 1e-5, torch.cuda.current_stream().cuda_stream)
    assert st == 0, lib.dafne_last_error()
    ref = F.relu(F.group_norm(x.float(), 32, gamma, beta, 1e-5)).permute(0, 2, 3, 1)
    assert torch.allclose(out.float(), ref, rtol=2e-3, atol=2e-3)


TAIL_CASES = [
    # name, N, H, W, K1, N1, N2: the bottleneck tails of res2 / res3 / res4, ragged sizes, several tiles per CTA
    ("res2_64_256_64", 2, 64, 64, 64, 256, 64),
    ("res3_128_512_128", 2, 48, 40, 128, 512, 128),
    ("res4_256_1024_256", 3, 32, 32, 256, 1024, 256),
    ("res4_odd_25x19", 2, 25, 19, 256, 1024, 256),
    ("res4_multiwave", 8, 64, 64, 256, 1024, 256),
    ("res3_multiwave", 2, 128, 128, 128, 512, 128),
    ("res2_small_7x5_n3", 3, 7, 5, 64, 256, 64),
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", TAIL_CASES, ids=[c[0] for c in TAIL_CASES])
def test_bottleneck_tail_matches_two_torch_convs(case):
    """csrc/tail_tc.cu: conv3 + BN + shortcut + ReLU and the next block's conv1 + BN + ReLU in one two-GEMM launch, against
    torch fp32 on the same fp16 inputs -- `out` rounded to fp16 before it feeds the second product (as through HBM)."""
    from dafne_b200 import _capi

    lib = _capi.lib()
    name, N, H, W, K1, N1, N2 = case
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(len(name))
    x = torch.randn(N, K1, H, W, generator=g).half()
    w3 = (torch.randn(N1, K1, 1, 1, generator=g) / K1 ** 0.5).half()
    w1 = (torch.randn(N2, N1, 1, 1, generator=g) / N1 ** 0.5).half()
    s1, b1 = torch.rand(N1, generator=g) * 0.5 + 0.25, torch.randn(N1, generator=g) * 0.5
    s2, b2 = torch.rand(N2, generator=g) + 0.5, torch.randn(N2, generator=g) * 0.5
    res = torch.randn(N, N1, H, W, generator=g).half()

    def nhwc(t):
        return t.to(dev).permute(0, 2, 3, 1).contiguous()

    x_d, res_d = nhwc(x), nhwc(res)
    w3_d, w1_d = w3.to(dev).reshape(N1, K1).contiguous(), w1.to(dev).reshape(N2, N1).contiguous()
    s1_d, b1_d, s2_d, b2_d = (t.to(dev).contiguous() for t in (s1, b1, s2, b2))
    out_d = torch.full((N, H, W, N1), float("nan"), device=dev, dtype=torch.float16)
    mid_d = torch.full((N, H, W, N2), float("nan"), device=dev, dtype=torch.float16)
    vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    _capi.check(lib.dafne_bottleneck_tail_nhwc(vp(x_d), N, H, W, K1, vp(w3_d), N1, vp(s1_d), vp(b1_d), vp(res_d),
                                               vp(out_d), vp(w1_d), N2, vp(s2_d), vp(b2_d), vp(mid_d),
                                               C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                "dafne_bottleneck_tail_nhwc")
    torch.cuda.synchronize()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    xr, rr = x.to(dev).float(), res.to(dev).float()
    ref_out = F.relu(F.conv2d(xr, w3.to(dev).float()) * s1_d.view(1, -1, 1, 1) + b1_d.view(1, -1, 1, 1) + rr)
    got_out = out_d.permute(0, 3, 1, 2).float()
    assert torch.isfinite(got_out).all()
    err = (got_out - ref_out).abs()
    assert bool((err <= 2e-3 * ref_out.abs().max() + 2e-3 * ref_out.abs()).all()), f"out: max err {err.max().item()}"
    # the second product sees the fp16-rounded block output the kernel itself wrote
    ref_mid = F.relu(F.conv2d(got_out, w1.to(dev).float()) * s2_d.view(1, -1, 1, 1) + b2_d.view(1, -1, 1, 1))
    got_mid = mid_d.permute(0, 3, 1, 2).float()
    assert torch.isfinite(got_mid).all()
    err = (got_mid - ref_mid).abs()
    assert bool((err <= 2e-3 * ref_mid.abs().max() + 2e-3 * ref_mid.abs()).all()), f"mid: max err {err.max().item()}"


GN_IN_CASES = [
    # name, N, H, W: the tower convolutions' shapes -- P3-like (several tiles per CTA, every tile another image), ragged
    # sizes whose halo crosses the image border on every side, a single tiny level
    ("p3_128x128_n4", 4, 128, 128),
    ("ragged_50x37_n3", 3, 50, 37),
    ("p7_8x8_n2", 2, 8, 8),
    ("wide_16x200_n1", 1, 16, 200),
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", GN_IN_CASES, ids=[c[0] for c in GN_IN_CASES])
def test_tower_conv_groupnorm_on_load(case):
    """conv_tc.cu mode 5: the tower convolution reads the previous layer's RAW output and applies that layer's
    GroupNorm(32) + ReLU to the landed halo box in shared memory (fcos/dafne head towers: Conv - GN - ReLU x 4). Against
    (a) torch fp32 group_norm + relu + conv2d on the same fp16 raw tensor and (b) this library's two-pass path (separate
    GroupNorm kernel, then the same convolution), whose only difference is the rounding of the affine form."""
    from dafne_b200 import _capi

    lib = _capi.lib()
    name, N, H, W = case
    Cc = 256
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(7 + len(name))
    x0 = torch.randn(N, Cc, H, W, generator=g).half()
    w0 = (torch.randn(Cc, Cc, 3, 3, generator=g) / (9 * Cc) ** 0.5).half()
    w1 = (torch.randn(Cc, Cc, 3, 3, generator=g) / (9 * Cc) ** 0.5).half()
    b0, b1 = torch.randn(Cc, generator=g) * 0.1, torch.randn(Cc, generator=g) * 0.1
    gamma, beta = torch.rand(Cc, generator=g) + 0.5, torch.randn(Cc, generator=g) * 0.3
    vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def nhwc(t):
        return t.to(dev).permute(0, 2, 3, 1).contiguous()

    def wk(t):  # [Cout][k*k][Cin]
        return t.to(dev).permute(0, 2, 3, 1).reshape(Cc, 9, Cc).contiguous()

    x_d, w0_d, w1_d = nhwc(x0), wk(w0), wk(w1)
    b0_d, b1_d, gamma_d, beta_d = (t.to(dev).contiguous() for t in (b0, b1, gamma, beta))
    # layer 0: raw output + statistics
    raw = torch.empty((N, H, W, Cc), device=dev, dtype=torch.float16)
    sums0 = torch.zeros((N, 32, 2), device=dev, dtype=torch.int64)
    _capi.check(lib.dafne_conv_nhwc(vp(x_d), N, H, W, Cc, vp(w0_d), Cc, 3, 1, None, vp(b0_d), 0, None, 0, 0, 0,
                                    vp(sums0), vp(raw), None, 0, stream), "dafne_conv_nhwc")
    # layer 1, GroupNorm on load
    out = torch.full((N, H, W, Cc), float("nan"), device=dev, dtype=torch.float16)
    sums1 = torch.zeros((N, 32, 2), device=dev, dtype=torch.int64)
    _capi.check(lib.dafne_conv_gn_in_nhwc(vp(raw), N, H, W, Cc, vp(sums0), vp(gamma_d), vp(beta_d), vp(w1_d), Cc,
                                          vp(b1_d), vp(sums1), vp(out), stream), "dafne_conv_gn_in_nhwc")
    # two-pass path of the same library
    normed = raw.clone()
    _capi.check(lib.dafne_gn_relu_nhwc(vp(normed), vp(normed), N, H * W, Cc, 32, vp(sums0), vp(gamma_d), vp(beta_d),
                                       C.c_float(1e-5), stream), "dafne_gn_relu_nhwc")
    out2 = torch.empty_like(out)
    sums2 = torch.zeros_like(sums1)
    _capi.check(lib.dafne_conv_nhwc(vp(normed), N, H, W, Cc, vp(w1_d), Cc, 3, 1, None, vp(b1_d), 0, None, 0, 0, 0,
                                    vp(sums2), vp(out2), None, 0, stream), "dafne_conv_nhwc")
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    raw_f = raw.permute(0, 3, 1, 2).float()
    ref_in = F.relu(F.group_norm(raw_f, 32, gamma_d, beta_d, 1e-5))
    ref = F.conv2d(ref_in.half().float(), w1.to(dev).float(), b1_d, padding=1)
    got = out.permute(0, 3, 1, 2).float()
    err = (got - ref).abs()
    assert bool((err <= 3e-3 * ref.abs().max() + 3e-3 * ref.abs()).all()), f"vs torch: max err {err.max().item()}"
    # against the two-pass path: the normalised input differs by at most one fp16 ulp on a few elements
    d2 = (out.float() - out2.float()).abs()
    assert d2.max().item() <= 4e-3 * out2.float().abs().max().item(), f"vs two-pass: {d2.max().item()}"
    assert (d2 > 0).float().mean().item() < 0.05
    # statistics of the output: fixed point, so equal wherever the outputs are equal; close otherwise
    rel = (sums1 - sums2).abs().double() / sums2.abs().double().clamp_min(1.0)
    assert rel.max().item() < 1e-3


PAIR_CASES = [
    # name, M, K, N, residual, relu: conv1 / conv3 of res4 and res5, ragged M (a pair tile whose second CTA is partly or
    # wholly outside), more tiles than pairs, one tile
    ("res4_conv1_1024_256", 3 * 32 * 32, 1024, 256, False, True),
    ("res4_conv3_256_1024_res", 3 * 32 * 32, 256, 1024, True, True),
    ("res5_conv3_512_2048_res", 2 * 16 * 16, 512, 2048, True, True),
    ("lateral_2048_256_linear", 2 * 16 * 16, 2048, 256, False, False),
    ("ragged_m_1000", 1000, 256, 512, True, True),
    ("ragged_m_130_one_cta_empty", 130, 128, 256, True, True),
    ("tiny_m_7", 7, 64, 256, False, True),
    ("multiwave_conv3", 8 * 64 * 64, 256, 1024, True, True),
    ("multiwave_conv1", 8 * 64 * 64, 1024, 256, False, True),
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", PAIR_CASES, ids=[c[0] for c in PAIR_CASES])
def test_pair_kernel_matches_torch(case):
    """csrc/pair_tc.cu: 1x1 convolutions as 256 x 256 tiles over a CTA pair (tcgen05.mma.cta_group::2), against torch fp32
    on the same fp16 inputs."""
    from dafne_b200 import _capi

    lib = _capi.lib()
    name, M, K, N, has_res, relu = case
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(11 + len(name))
    x = torch.randn(M, K, generator=g).half().to(dev)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).half().to(dev)
    sc = (torch.rand(N, generator=g) + 0.5).to(dev)
    sh = (torch.randn(N, generator=g) * 0.5).to(dev)
    res = torch.randn(M, N, generator=g).half().to(dev) if has_res else None
    out = torch.full((M, N), float("nan"), device=dev, dtype=torch.float16)
    vp = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731
    _capi.check(lib.dafne_conv1x1_pair_nhwc(vp(x), M, K, vp(w), N, vp(sc), vp(sh), int(relu), vp(res), vp(out),
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                "dafne_conv1x1_pair_nhwc")
    torch.cuda.synchronize()
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = (x.float() @ w.float().t()) * sc + sh
    if has_res:
        ref = ref + res.float()
    if relu:
        ref = F.relu(ref)
    got = out.float()
    assert torch.isfinite(got).all()
    err = (got - ref).abs()
    assert bool((err <= 2e-3 * ref.abs().max() + 2e-3 * ref.abs()).all()), f"max err {err.max().item()}"
