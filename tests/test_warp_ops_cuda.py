"""GPU: the feature-warp / bilinear-resize / PReLU / pose-matrix kernels (csrc/warp_cl.cu, through the C ABI) against the torch
operators the reference calls at those sites -- F.grid_sample inside IFRNet.warp (IFRNet.py:7-15), F.interpolate
(hrnet_encoder.py:275-280, IFRNet.py:118, fusion_module.py:68-99), nn.PReLU, transformation_from_parameters (layers.py:28-103) --
evaluated in float64 on the same inputs.  Tolerances: 2e-6 of the tensor scale forward (fp32 rounding of a 4-term blend),
1e-5 backward; the deterministic scatter must also be bitwise repeatable."""
import pytest

pytestmark = pytest.mark.gpu


def _cl(t):
    import torch
    return t.contiguous(memory_format=torch.channels_last)


def _ref_warp(img, flow):
    import torch
    import torch.nn.functional as F
    B, _, H, W = flow.shape
    xx = torch.linspace(-1.0, 1.0, W, device=flow.device, dtype=flow.dtype).view(1, 1, 1, W).expand(B, -1, H, -1)
    yy = torch.linspace(-1.0, 1.0, H, device=flow.device, dtype=flow.dtype).view(1, 1, H, 1).expand(B, -1, -1, W)
    grid = torch.cat([xx + flow[:, 0:1] / ((W - 1.0) / 2.0), yy + flow[:, 1:2] / ((H - 1.0) / 2.0)], 1).to(img)
    return F.grid_sample(img, grid.permute(0, 2, 3, 1), mode="bilinear", padding_mode="border", align_corners=True)


@pytest.mark.parametrize("B,C,H,W,cl", [(2, 64, 24, 40, True), (1, 18, 17, 23, True), (3, 3, 32, 48, False), (2, 512, 6, 20, True)])
def test_flow_warp_forward_backward(B, C, H, W, cl):
    import torch
    from mono_vifi_b200 import warp_ops
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(5)
    x = torch.randn(B, C, H, W, device=dev, generator=g)
    flow = torch.randn(B, 2, H, W, device=dev, generator=g) * 6.0   # a good share of the samples leaves the frame (border clamp)
    xin = (_cl(x) if cl else x).requires_grad_(cl)
    y = warp_ops.flow_warp(xin, flow)
    xd = x.double().requires_grad_(True)
    yr = _ref_warp(xd, flow.double())
    assert float((y.double() - yr).abs().max()) <= 2e-5 * float(yr.abs().max())
    if cl:
        gy = torch.randn(B, C, H, W, device=dev, generator=g)
        y.backward(gy)
        yr.backward(gy.double())
        assert float((xin.grad.double() - xd.grad).abs().max()) <= 1e-5 * float(xd.grad.abs().max())
        g1 = xin.grad.clone()
        xin.grad = None
        warp_ops.flow_warp(xin, flow).backward(gy)
        assert torch.equal(g1, xin.grad)   # fixed-point scatter: bitwise repeatable


@pytest.mark.parametrize("C,Hi,Wi,size,sf,align,cl", [
    (36, 12, 20, (48, 80), None, True, True),      # HRNet fuse: x4, align_corners=True
    (18, 6, 10, (48, 80), None, True, True),       # x8, 18 channels (float2 path)
    (24, 20, 64, None, 2.0, False, True),          # Lite-Mono decoder: x2, align_corners=False
    (2, 48, 80, None, 0.5, False, False),          # flow pyramid of the fusion module
    (2, 24, 40, (37, 59), None, False, False),     # odd ratio
    (1, 192, 320, (192, 640), None, False, False),  # IFRNet mask to full resolution
    (3, 320, 1024, (192, 320), None, False, False),  # IFRNet input resize at 320x1024
])
def test_resize_bilinear_forward_backward(C, Hi, Wi, size, sf, align, cl):
    import torch
    import torch.nn.functional as F
    from mono_vifi_b200 import warp_ops
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(7)
    x = torch.randn(2, C, Hi, Wi, device=dev, generator=g)
    mul = (1.0, 1.0) if cl else (1.75, 0.5)
    xin = (_cl(x) if cl else x.clone()).requires_grad_(True)
    y = warp_ops.resize_bilinear(xin, size=size, scale_factor=sf, align_corners=align, mul=mul)
    xd = x.double().requires_grad_(True)
    yr = F.interpolate(xd, size=size, scale_factor=sf, mode="bilinear", align_corners=align)
    yr = yr * torch.tensor([mul[i & 1] for i in range(C)], device=dev, dtype=torch.float64).view(1, -1, 1, 1)
    assert y.shape == yr.shape
    # (the fp32 source coordinate carries ~1e-4 px of rounding at x ~ 1000: the 320x1024 case)
    assert float((y.double() - yr).abs().max()) <= 1e-4 * float(yr.abs().max())
    gy = torch.randn(y.shape, device=dev, generator=g)
    y.backward(gy)
    yr.backward(gy.double())
    assert float((xin.grad.double() - xd.grad).abs().max()) <= 1e-4 * float(xd.grad.abs().max())


def test_prelu_tail():
    import torch
    import torch.nn.functional as F
    from mono_vifi_b200 import warp_ops
    dev = torch.device("cuda:0")
    x, r = torch.randn(2, 72, 12, 20, device=dev), torch.randn(2, 72, 12, 20, device=dev)
    s = torch.rand(72, device=dev)
    with torch.no_grad():
        n0 = warp_ops.launches["prelu"]
        assert torch.equal(warp_ops.prelu(_cl(x), s, _cl(r)), F.prelu(x + r, s))
        assert torch.equal(warp_ops.prelu(x, s), F.prelu(x, s))
        assert warp_ops.launches["prelu"] == n0 + 2


@pytest.mark.parametrize("invert", [False, True])
def test_pose_matrix_matches_the_op_by_op_form(invert):
    import torch
    from mono_vifi_b200 import layers as L, warp_ops
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(11)
    aa = (0.3 * torch.randn(12, 1, 3, generator=g)).to(dev)
    tr = (0.5 * torch.randn(12, 1, 3, generator=g)).to(dev)
    a1, t1 = aa.clone().requires_grad_(True), tr.clone().requires_grad_(True)
    M = warp_ops.pose_matrix(a1, t1, invert)
    a2, t2 = aa.double().cpu().requires_grad_(True), tr.double().cpu().requires_grad_(True)
    Mr = L.transformation_from_parameters(a2, t2, invert)   # CPU tensors take the reference's op sequence
    assert float((M.double().cpu() - Mr).abs().max()) <= 2e-6
    gM = torch.randn(12, 4, 4, generator=g)
    M.backward(gM.to(dev))
    Mr.backward(gM.double())
    assert float((a1.grad.double().cpu() - a2.grad).abs().max()) <= 1e-5 * float(a2.grad.abs().max())
    assert float((t1.grad.double().cpu() - t2.grad).abs().max()) <= 1e-5 * float(t2.grad.abs().max())
    # the layers.* entry point takes the kernel on CUDA tensors
    n0 = warp_ops.launches["pose_matrix"]
    assert torch.equal(L.transformation_from_parameters(aa, tr, invert), M.detach())
    assert warp_ops.launches["pose_matrix"] == n0 + 1


@pytest.mark.parametrize("cin,cout,H,W", [(384, 148, 6, 10), (162, 58, 12, 20), (192, 8, 24, 40)])
def test_transposed_convolution_on_the_stride2_dgrad_kernel(cin, cout, H, W):
    """nn.ConvTranspose2d(cin, cout, 4, 2, 1) of IFRNet's decoders (IFRNet.py:194) vs torch fp64; TF32 operand rounding: 2e-3"""
    import torch
    import torch.nn.functional as F
    from mono_vifi_b200 import conv_tc
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(3)
    x = torch.randn(2, cin, H, W, device=dev, generator=g)
    w = torch.randn(cin, cout, 4, 4, device=dev, generator=g) / (cin * 4) ** 0.5
    b = torch.randn(cout, device=dev, generator=g)
    with torch.no_grad():
        y = conv_tc.conv_transpose2d_s2_inference(x, w, b, 1)
    yr = F.conv_transpose2d(x.double(), w.double(), b.double(), stride=2, padding=1)
    assert y.shape == yr.shape
    assert float((y.double() - yr).abs().max()) <= 2e-3 * float(yr.abs().max())


@pytest.mark.parametrize("cin,cout,k,stride", [(64, 64, 3, 1), (24, 36, 3, 2), (436, 432, 3, 1), (4, 24, 3, 2)])
def test_convolution_with_prelu_epilogue(cin, cout, k, stride):
    import torch
    import torch.nn.functional as F
    from mono_vifi_b200 import conv_tc
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(4)
    x = torch.randn(2, cin, 24, 40, device=dev, generator=g)
    w = torch.randn(cout, cin, k, k, device=dev, generator=g) / (cin * k * k) ** 0.5
    b = torch.randn(cout, device=dev, generator=g)
    s = torch.rand(cout, device=dev, generator=g)
    with torch.no_grad():
        y = conv_tc.conv2d_prelu_inference(x, w, b, s, stride, k // 2)
    yr = F.prelu(F.conv2d(x.double(), w.double(), b.double(), stride, k // 2), s.double())
    assert float((y.double() - yr).abs().max()) <= 2e-3 * float(yr.abs().max())
