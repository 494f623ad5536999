"""GPU parity of the tcgen05 implicit-GEMM convolutions (through the C ABI) against torch's fp32 convolution
(TF32 disabled: a true-fp32 reference).  TF32 rounds each input to 10 mantissa bits, so the tolerance is
2e-3 of the output's magnitude (north_star: depth tensors within 1e-3 relative after the whole network)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ref(x, w, b, pad, stride=1):
    import torch.nn.functional as F
    return F.conv2d(x.double(), w.double(), None if b is None else b.double(), stride, pad).float()


CASES = [
    # B, Cin, H, W, Cout, k, pad, stride, bias
    (1, 8, 4, 32, 16, 1, 0, 1, False),      # one tile, a quarter-full stage (zero-filled channels)
    (1, 32, 4, 32, 16, 1, 0, 1, False),     # one full stage
    (1, 64, 8, 64, 32, 1, 0, 1, True),      # several tiles and stages
    (2, 64, 12, 40, 64, 3, 1, 1, True),     # 3x3 zero padding, ragged tile edges
    (2, 16, 9, 37, 16, 3, 1, 1, False),     # odd width, Cin 16
    (1, 96, 16, 64, 32, 3, 1, 1, True),     # Cin not a power of two
    (2, 128, 24, 80, 128, 3, 1, 1, False),  # ResNet layer2 shape
    (1, 256, 12, 40, 256, 3, 1, 1, False),  # two N tiles
    (2, 512, 6, 20, 512, 3, 1, 1, False),   # ResNet layer4 shape: four N tiles, 144 pipeline iterations
    (1, 64, 10, 36, 64, 3, 0, 1, True),     # valid convolution (reflection-padded input comes this way)
    (1, 64, 8, 32, 24, 3, 2, 1, False),     # "full" padding = what dgrad of a pad-0 conv uses; Cout not a multiple of 16
    (2, 16, 16, 64, 1, 3, 1, 1, True),      # dispconv: a single output channel
    (2, 64, 24, 80, 128, 3, 1, 2, False),   # stride 2 (ResNet layer2.0.conv1)
    (2, 64, 24, 80, 128, 1, 0, 2, False),   # 1x1 stride 2 (downsample)
    (1, 8, 33, 71, 64, 7, 3, 2, False),     # 7x7 stride 2 stem on odd sizes (image channels padded 3 -> 8)
    (1, 4, 16, 32, 16, 3, 1, 1, False),     # four input channels (16-byte pixels)
    (12, 64, 48, 160, 64, 3, 1, 1, True),   # ResNet layer1 at the benchmark size: 4 column segments, 2 stacked tiles per CTA
    (2, 16, 64, 640, 16, 3, 1, 1, False),   # full-resolution decoder layer: 6 column segments of 107
    (2, 32, 32, 322, 16, 3, 0, 1, True),    # reflection-padded input of a 320-wide layer (valid conv)
    (3, 128, 24, 80, 128, 3, 1, 1, False),  # 2 segments x 3 rows
    (2, 64, 7, 9, 32, 5, 2, 1, False),      # 5x5 filter, tiny image
    (1, 32, 5, 300, 16, 1, 0, 1, False),    # 1x1 on a wide image (per-tap kernel)
]


@pytest.mark.parametrize("case", CASES)
def test_forward_vs_fp32(case):
    import torch
    from mono_vifi_b200 import conv_tc
    B, Cin, H, W, Cout, k, pad, stride, bias = case
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(B, Cin, H, W, device="cuda", generator=g)
    w = torch.randn(Cout, Cin, k, k, device="cuda", generator=g) / (Cin * k * k) ** 0.5
    b = torch.randn(Cout, device="cuda", generator=g) if bias else None
    assert conv_tc.supported(x, w, stride, pad)
    y = conv_tc.conv_forward_raw(x, conv_tc.pack_filters(w), b, Cout, k, k, pad, stride)
    torch.cuda.synchronize()
    ref = _ref(x, w, b, pad, stride)
    assert y.shape == ref.shape
    err = (y - ref).abs().max().item()
    assert err <= 2e-3 * ref.abs().max().item(), (err, ref.abs().max().item())


def test_strided_input_and_output_views():
    """channel slices of a wider tensor (what torch.cat would otherwise copy) and a channel-offset output"""
    import torch
    from mono_vifi_b200 import conv_tc
    g = torch.Generator(device="cuda").manual_seed(6)
    big = torch.randn(2, 96, 12, 64, device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
    x = big[:, 32:96]
    w = torch.randn(32, 64, 3, 3, device="cuda", generator=g) / 24.0
    out = torch.zeros(2, 48, 12, 64, device="cuda").contiguous(memory_format=torch.channels_last)
    conv_tc.conv_forward_raw(x, conv_tc.pack_filters(w), None, 32, 3, 3, 1, out=out[:, 16:48])
    ref = _ref(x.contiguous(), w, None, 1)
    assert (out[:, 16:48] - ref).abs().max().item() <= 2e-3 * ref.abs().max().item()
    assert out[:, :16].abs().max().item() == 0.0


@pytest.mark.parametrize("case", [(2, 64, 12, 40, 64, 3, 1, 1), (1, 32, 8, 64, 16, 3, 1, 1), (2, 64, 10, 36, 32, 3, 0, 1),
                                  (1, 128, 6, 20, 64, 1, 0, 1), (2, 64, 12, 40, 128, 3, 1, 2), (1, 16, 9, 33, 1, 3, 1, 1)])
def test_autograd_vs_fp32(case):
    import torch
    from mono_vifi_b200 import conv_tc
    B, Cin, H, W, Cout, k, pad, stride = case
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(B, Cin, H, W, device="cuda", generator=g, requires_grad=True)
    w = (torch.randn(Cout, Cin, k, k, device="cuda", generator=g) / (Cin * k * k) ** 0.5).requires_grad_(True)
    b = torch.randn(Cout, device="cuda", generator=g, requires_grad=True)
    y = conv_tc.conv2d(x, w, b, stride, pad)
    gy = torch.randn(y.shape, device="cuda", generator=g)
    y.backward(gy)
    xr, wr, br = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    yr = torch.nn.functional.conv2d(xr, wr, br, stride, pad)
    yr.backward(gy.double())
    for got, ref in ((x.grad, xr.grad), (w.grad, wr.grad), (b.grad, br.grad)):
        ref = ref.float()
        assert (got - ref).abs().max().item() <= 3e-3 * ref.abs().max().item()


def test_unsupported_shapes_are_reported():
    import torch
    from mono_vifi_b200 import _lib, conv_tc
    x = torch.randn(1, 3, 8, 32, device="cuda")
    w = torch.randn(8, 3, 3, 3, device="cuda")
    assert not conv_tc.supported(x, w, 1, 1)           # Cin % 4
    assert not conv_tc.supported(torch.randn(1, 8, 8, 32, device="cuda"), torch.randn(8, 8, 3, 3, device="cuda"), 3, 1)
    d = _lib.Conv2dDesc(1, 3, 8, 32, 8, 3, 3, 1, 1)
    assert _lib.lib().mvf_conv2d_supported(d) == 0 and b"multiple of 4" in _lib.lib().mvf_last_error()


def test_umma_selftest_pins_descriptors():
    """D = A . B^T on one CTA for both operand layouts the kernels use (K-major; MN-major with the 32-byte-unit swizzle)"""
    import torch
    from mono_vifi_b200 import _lib
    L = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device="cuda").manual_seed(1)
    for N, K in ((16, 32), (64, 64), (256, 96)):
        # K-major A [160][K]; rows [off, off+128) are the operand (off != 0: start not aligned to the swizzle pattern)
        A = torch.randn(160, K, device="cuda", generator=g)
        Bm = torch.randn(N, K, device="cuda", generator=g)
        for off in (0, 1, 5, 17):
            D = torch.full((128, N), -5.0, device="cuda")
            _lib.check(L.mvf_selftest_umma_rows(A.data_ptr(), Bm.data_ptr(), D.data_ptr(), N, K, off, 0, st), "mvf_selftest_umma_rows")
            ref = A[off:off + 128].double() @ Bm.double().t()
            assert (D.double() - ref).abs().max().item() <= 2e-3 * ref.abs().max().item() * (K / 32) ** 0.5
    # MN-major A [K+8][128] (the layout NCHW / pixel-major operands have), k-rows shifted by off
    N, K = 64, 32
    A = torch.randn(K + 8, 128, device="cuda", generator=g)
    Bm = torch.randn(N, K, device="cuda", generator=g)
    for off in (0, 1, 3, 8):
        D = torch.full((128, N), -5.0, device="cuda")
        _lib.check(L.mvf_selftest_umma_rows(A.data_ptr(), Bm.data_ptr(), D.data_ptr(), N, K, off, 2, st), "mvf_selftest_umma_rows")
        ref = A[off:off + K].double().t() @ Bm.double().t()
        assert (D.double() - ref).abs().max().item() <= 2e-3 * ref.abs().max().item()


WGRAD_CASES = [
    # B, Cin, H, W, Cout, k, pad, stride
    (1, 32, 4, 32, 32, 1, 0, 1),      # one patch per image row group, one tap
    (2, 64, 12, 40, 64, 3, 1, 1),     # 3x3, ragged patches
    (2, 16, 16, 64, 16, 3, 1, 1),     # narrow channels (zero-filled cout / cin blocks)
    (1, 96, 16, 64, 32, 3, 1, 1),     # three cin tiles
    (2, 128, 24, 80, 128, 3, 1, 1),   # ResNet layer2
    (2, 512, 6, 20, 512, 3, 1, 1),    # ResNet layer4: 4 x 16 tiles, few pixels
    (2, 64, 24, 80, 128, 3, 1, 2),    # stride 2
    (2, 64, 24, 80, 128, 1, 0, 2),    # 1x1 stride 2 (downsample)
    (2, 64, 10, 36, 32, 3, 0, 1),     # valid convolution
    (3, 256, 6, 20, 12, 1, 0, 1),     # pose head: 12 output channels
    (12, 64, 48, 160, 64, 3, 1, 1),   # ResNet layer1 at the benchmark size (split-K over 148 CTAs)
    (2, 36, 24, 80, 36, 3, 1, 1),     # HRNet-W18 branch widths: Cout above 32 and not a multiple of 32 (per-block boxes, zero fill)
    (2, 72, 12, 40, 72, 3, 1, 1),
    (2, 144, 6, 20, 144, 3, 1, 1),
    (2, 36, 24, 80, 72, 3, 1, 2),     # HRNet transition: stride 2, 36 -> 72
    (2, 144, 6, 20, 36, 1, 0, 1),     # HRNet fuse layer: 1x1, 144 -> 36
    (2, 20, 48, 160, 20, 3, 1, 1),    # the 18-channel branch as the layers run it (zero-padded to 20)
]


@pytest.mark.parametrize("case", WGRAD_CASES)
def test_wgrad_vs_fp64(case):
    import torch
    from mono_vifi_b200 import conv_tc
    B, Cin, H, W, Cout, k, pad, stride = case
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(B, Cin, H, W, device="cuda", generator=g)
    Ho, Wo = conv_tc.out_hw(H, W, k, k, pad, stride)
    gy = torch.randn(B, Cout, Ho, Wo, device="cuda", generator=g)
    before = dict(conv_tc.launches)
    gw = conv_tc.weight_grad(conv_tc._as_input(x), gy, (Cout, Cin, k, k), pad, stride)
    torch.cuda.synchronize()
    assert conv_tc.launches["wgrad"] == before["wgrad"] + 1, "the tcgen05 wgrad kernel must cover this shape"
    wr = torch.zeros(Cout, Cin, k, k, device="cuda", dtype=torch.float64, requires_grad=True)
    torch.nn.functional.conv2d(x.double(), wr, None, stride, pad).backward(gy.double())
    ref = wr.grad.float()
    err = (gw - ref).abs().max().item()
    assert err <= 3e-3 * ref.abs().max().item(), (err, ref.abs().max().item())
    # bitwise reproducible (fixed-order split-K reduction)
    gw2 = conv_tc.weight_grad(conv_tc._as_input(x), gy, (Cout, Cin, k, k), pad, stride)
    assert torch.equal(gw, gw2)


@pytest.mark.parametrize("cin", [3, 6])
def test_row_packed_stem_matches_fp64(cin):
    """7x7 stride-2 pad-3 stem on 3- / 6-channel images through the row-packed path (fprop + wgrad on tcgen05)"""
    import torch
    from mono_vifi_b200 import conv, conv_tc
    conv.set_backend("tcgen05")
    g = torch.Generator(device="cuda").manual_seed(21)
    x = torch.randn(2, cin, 64, 96, device="cuda", generator=g)
    w = (torch.randn(64, cin, 7, 7, device="cuda", generator=g) / (cin * 49) ** 0.5).requires_grad_(True)
    before = dict(conv_tc.launches)
    y = conv.conv2d(x, w, None, 2, 3)
    gy = torch.randn(y.shape, device="cuda", generator=g)
    y.backward(gy)
    assert conv_tc.launches["fprop"] == before["fprop"] + 1 and conv_tc.launches["wgrad"] == before["wgrad"] + 1
    wr = w.detach().double().requires_grad_(True)
    yr = torch.nn.functional.conv2d(x.double(), wr, None, 2, 3)
    yr.backward(gy.double())
    assert y.shape == yr.shape
    assert (y.double() - yr).abs().max().item() <= 2e-3 * yr.abs().max().item()
    assert (w.grad.double() - wr.grad).abs().max().item() <= 3e-3 * wr.grad.abs().max().item()


DGRAD_S2_CASES = [
    # B, Cin, H, W, Cout, k, pad
    (2, 64, 24, 80, 128, 3, 1),      # ResNet layer2.0.conv1
    (2, 64, 24, 80, 128, 1, 0),      # its 1x1 shortcut (three empty parity classes)
    (1, 32, 9, 13, 16, 3, 1),        # odd sizes: ragged class lattices
    (2, 256, 12, 40, 512, 3, 1),     # layer4.0.conv1: several k blocks, 4 N tiles
    (1, 16, 10, 10, 16, 5, 2),
    (2, 4, 32, 48, 64, 7, 3),        # the stem (its data gradient is only needed by the pose net's callers' tests)
]


@pytest.mark.parametrize("case", DGRAD_S2_CASES)
def test_dgrad_stride2_vs_fp64(case):
    import torch
    from mono_vifi_b200 import conv_tc
    B, Cin, H, W, Cout, k, pad = case
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(B, Cin, H, W, device="cuda", generator=g)
    w = torch.randn(Cout, Cin, k, k, device="cuda", generator=g) / (Cin * k * k) ** 0.5
    Ho, Wo = conv_tc.out_hw(H, W, k, k, pad, 2)
    gy = torch.randn(B, Cout, Ho, Wo, device="cuda", generator=g)
    gx = conv_tc.input_grad_s2(conv_tc._as_input(x), gy, w, pad)
    torch.cuda.synchronize()
    xr = x.double().requires_grad_(True)
    torch.nn.functional.conv2d(xr, w.double(), None, 2, pad).backward(gy.double())
    ref = xr.grad.float()
    err = (gx - ref).abs().max().item()
    assert err <= 3e-3 * ref.abs().max().item(), (err, ref.abs().max().item())


@pytest.mark.parametrize("case", [(2, 16, 16, 64), (1, 64, 9, 33), (3, 32, 7, 5), (12, 16, 192, 640)])
def test_dispconv_direct_kernels_vs_fp64(case):
    """Conv3x3(C, 1) on a padded channels-last feature (csrc/dispconv.cu): fp32 FMA chains, so 1e-5 of the magnitude; the weight
    gradient's split reduction has a fixed order (bitwise repeatable)."""
    import torch
    from mono_vifi_b200 import conv_tc
    B, C, H, W = case
    g = torch.Generator(device="cuda").manual_seed(11)
    xp = torch.randn(B, C, H + 2, W + 2, device="cuda", generator=g).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    w = (torch.randn(1, C, 3, 3, device="cuda", generator=g) / (C * 9) ** 0.5).requires_grad_(True)
    b = torch.randn(1, device="cuda", generator=g, requires_grad=True)
    assert conv_tc.dispconv_supported(xp, w, 1, 0)
    n0 = dict(conv_tc.launches)
    y = conv_tc.conv2d(xp, w, b, 1, 0)
    gy = torch.randn(y.shape, device="cuda", generator=g)
    y.backward(gy)
    assert conv_tc.launches["fprop"] == n0["fprop"] + 1
    xr, wr, br = (t.detach().double().requires_grad_(True) for t in (xp, w, b))
    yr = torch.nn.functional.conv2d(xr, wr, br)
    yr.backward(gy.double())
    for got, ref in ((y, yr), (xp.grad, xr.grad), (w.grad, wr.grad), (b.grad, br.grad)):
        ref = ref.float()
        assert got.shape == ref.shape
        assert (got - ref).abs().max().item() <= 2e-5 * ref.abs().max().item() + 1e-6, (got - ref).abs().max().item()
    gw1 = w.grad.clone()
    w.grad = None
    xp.grad = None
    conv_tc.conv2d(xp, w, b, 1, 0).backward(gy)
    assert torch.equal(gw1, w.grad)


@pytest.mark.parametrize("shape", [(2, 64, 12, 40, 64, 3, 1), (2, 512, 2, 3, 512, 3, 1), (6, 512, 2, 3, 512, 3, 1), (3, 128, 24, 80, 128, 3, 1),
                                   (2, 16, 16, 64, 16, 3, 1), (2, 256, 6, 20, 256, 3, 0), (12, 512, 2, 3, 256, 3, 1)])
def test_every_tile_shape_of_the_patch_kernel(shape):
    """The stride-1 kernel's tile shape (N_TILE x MT stacked tiles; MT = 1 splits the taps over the two MMA issuers) is picked by a cost
    model; every shape it can pick must give the same convolution (MVF_CONV_TILE_RULE forces one)."""
    import torch
    from mono_vifi_b200 import conv_tc
    B, Cin, H, W, Cout, k, pad = shape
    g = torch.Generator(device="cuda").manual_seed(23)
    x = torch.randn(B, H, W, Cin, device="cuda", generator=g).permute(0, 3, 1, 2)
    w = torch.randn(Cout, Cin, k, k, device="cuda", generator=g) / (Cin * k * k) ** 0.5
    b = torch.randn(Cout, device="cuda", generator=g)
    ref = _ref(x, w, b, pad)
    wp = conv_tc.pack_filters(w)
    n_wide = 16
    while n_wide < Cout and n_wide < 128:
        n_wide *= 2
    old = os.environ.get("MVF_CONV_TILE_RULE")
    try:
        outs = {}
        for rule in [None, "r1"] + ["%d,%d" % (n, mt) for n in (16, 32, 64, 128) if n <= n_wide for mt in (1, 2)]:
            if rule is None:
                os.environ.pop("MVF_CONV_TILE_RULE", None)
            else:
                os.environ["MVF_CONV_TILE_RULE"] = rule
            for rep in range(3):   # repeated launches: the issuers' barrier phases wrap differently from call to call
                y = conv_tc.conv_forward_raw(x, wp, b, Cout, k, k, pad, 1)
            torch.cuda.synchronize()
            err = (y - ref).abs().max().item()
            assert err <= 2e-3 * ref.abs().max().item(), (rule, err, ref.abs().max().item())
            outs[rule] = y
    finally:
        if old is None:
            os.environ.pop("MVF_CONV_TILE_RULE", None)
        else:
            os.environ["MVF_CONV_TILE_RULE"] = old


def test_shapes_outside_the_kernels_raise_instead_of_reaching_a_library():
    """The product path has no silent (nor default) library fallback: a CUDA convolution the tcgen05 kernels do not cover raises;
    MVF_LIBRARY_FALLBACK=1 (conv.library_fallback) turns it into a counted cuDNN call."""
    import torch
    from mono_vifi_b200 import conv
    x = torch.randn(2, 8, 12, 16, device="cuda")
    w = torch.randn(8, 4, 3, 3, device="cuda")          # groups = 2
    assert conv.get_backend() == "tcgen05" and not conv.library_fallback
    with pytest.raises(conv.UnsupportedConvolution):
        conv.conv2d(x, w, None, 1, 1, 1, 2)
    with pytest.raises(conv.UnsupportedConvolution):
        conv.conv2d(x, torch.randn(8, 8, 3, 3, device="cuda"), None, 1, 2, 2, 1)   # dilation 2
    n0 = conv.stats["cudnn"]
    conv.library_fallback = True
    try:
        y = conv.conv2d(x, w, None, 1, 1, 1, 2)
    finally:
        conv.library_fallback = False
    assert conv.stats["cudnn"] == n0 + 1
    assert torch.allclose(y, torch.nn.functional.conv2d(x, w, None, 1, 1, 1, 2))
