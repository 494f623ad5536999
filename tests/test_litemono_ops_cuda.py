"""GPU: the Lite-Mono block kernels (csrc/litemono.cu, through the C ABI) against the torch operators the reference calls --
nn.Conv2d(groups=C, dilation=d) (LiteMono.py:140-155), nn.GELU, F.layer_norm (LiteMono.py:93-121) -- evaluated in float64.
Tolerance 1e-5 of the tensor scale (plain fp32 arithmetic, no tensor cores involved)."""
import pytest

pytestmark = pytest.mark.gpu


def _close(a, b, tol=1e-5):
    return float((a.double() - b).abs().max()) <= tol * max(float(b.abs().max()), 1e-30)


@pytest.mark.parametrize("B,C,H,W,d", [(2, 48, 20, 64, 1), (2, 80, 10, 32, 2), (1, 128, 20, 64, 5), (2, 128, 5, 16, 10), (3, 224, 12, 40, 3)])
def test_depthwise_dilated_conv(B, C, H, W, d):
    import torch
    import torch.nn.functional as F
    from mono_vifi_b200 import litemono_ops as O
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(2)
    x = torch.randn(B, C, H, W, device=dev, generator=g)
    w = torch.randn(C, 1, 3, 3, device=dev, generator=g)
    x1 = x.contiguous(memory_format=torch.channels_last).requires_grad_(True)
    w1 = w.clone().requires_grad_(True)
    y = O.dwconv3x3(x1, w1, None, d)
    xd, wd = x.double().requires_grad_(True), w.double().requires_grad_(True)
    yr = F.conv2d(xd, wd, None, 1, d, d, C)
    assert _close(y, yr)
    gy = torch.randn(y.shape, device=dev, generator=g)
    y.backward(gy)
    yr.backward(gy.double())
    assert _close(x1.grad, xd.grad) and _close(w1.grad, wd.grad, 2e-5)
    g1 = w1.grad.clone()
    w1.grad = None
    O.dwconv3x3(x1, w1, None, d).backward(gy)
    assert torch.equal(g1, w1.grad)   # fixed-order reduction


def test_gelu():
    import torch
    import torch.nn.functional as F
    from mono_vifi_b200 import litemono_ops as O
    dev = torch.device("cuda:0")
    x = (3.0 * torch.randn(2, 288, 20, 64, device=dev)).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    y = O.gelu(x)
    xd = x.detach().double().requires_grad_(True)
    yr = F.gelu(xd)
    assert y.stride() == x.stride() and _close(y, yr, 2e-6)
    gy = torch.randn_like(y)
    y.backward(gy)
    yr.backward(gy.double())
    assert _close(x.grad, xd.grad, 2e-6)


@pytest.mark.parametrize("C,P", [(48, (2, 20, 64)), (128, (3, 5 * 16)), (224, (2, 7, 9))])
def test_layernorm_channels_last(C, P):
    import torch
    import torch.nn.functional as F
    from mono_vifi_b200 import litemono_ops as O
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(6)
    x = (torch.randn(*P, C, device=dev, generator=g) * 2 + 0.5).requires_grad_(True)
    w = torch.randn(C, device=dev, generator=g).requires_grad_(True)
    b = torch.randn(C, device=dev, generator=g).requires_grad_(True)
    y = O.layer_norm_cl(x, w, b, 1e-6)
    xd, wd, bd = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    yr = F.layer_norm(xd, (C,), wd, bd, 1e-6)
    assert _close(y, yr)
    gy = torch.randn(y.shape, device=dev, generator=g)
    y.backward(gy)
    yr.backward(gy.double())
    assert _close(x.grad, xd.grad, 2e-5) and _close(w.grad, wd.grad, 2e-5) and _close(b.grad, bd.grad, 2e-5)


def test_litemono_blocks_use_the_kernels():
    import torch
    from mono_vifi_b200 import conv, litemono_ops as O, networks as N
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    enc = N.LiteMono.DepthEncoder(model="lite-mono", drop_path_rate=0.0, width=640, height=192).to(dev).train()
    n0, c0 = dict(O.launches), dict(conv.stats)
    feats = enc(torch.rand(2, 3, 192, 640, device=dev))
    sum(f.mean() for f in feats).backward()
    assert O.launches["dwconv_fwd"] - n0["dwconv_fwd"] == 15 and O.launches["dwconv_wgrad"] - n0["dwconv_wgrad"] == 15
    assert O.launches["ln_fwd"] - n0["ln_fwd"] == 6 and O.launches["gelu_bwd"] - n0["gelu_bwd"] >= 18
    assert conv.stats["cudnn"] == c0["cudnn"], "no convolution of the encoder may fall back to the library"
