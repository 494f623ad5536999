"""Host-side plan of the stride-2 data-gradient kernel (mvf_conv2d_dgrad_s2_plan, csrc/conv_tc.cu): the parity classes,
tap offsets and packed-bank tap indices the kernel walks, checked on the CPU by evaluating exactly that table with
torch and comparing with autograd's input gradient of F.conv2d(stride=2) -- the backward of the down-sampling
convolutions of the ResNet encoders (networks/monodepth2.py:16-31).  No GPU compute: the library only has to load."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from mono_vifi_b200 import _lib


def plan(KH, KW, pad):
    buf = (ctypes.c_int * 512)()
    n = _lib.lib().mvf_conv2d_dgrad_s2_plan(KH, KW, pad, buf, 512)
    assert 0 < n <= 512
    t, k, classes = list(buf[:n]), 1, []
    for _ in range(t[0]):
        py, px, ntaps = t[k:k + 3]
        k += 3
        taps = [tuple(t[k + 3 * i:k + 3 * i + 3]) for i in range(ntaps)]
        k += 3 * ntaps
        classes.append((py, px, taps))
    assert k == n
    return classes


@pytest.mark.parametrize("case", [(3, 3, 1, 10, 14), (3, 3, 1, 9, 13), (1, 1, 0, 8, 12), (1, 1, 0, 7, 9), (7, 7, 3, 12, 16),
                                  (3, 3, 0, 9, 11), (5, 5, 2, 10, 10), (3, 1, 1, 8, 8)])
def test_plan_reproduces_the_input_gradient(case):
    KH, KW, pad, H, W = case
    pad_w = pad if KW > 1 else 0
    if KH != KW:                      # the kernel has one `pad`; rectangular filters are only planned for equal padding
        pad_w = pad
    torch.manual_seed(0)
    B, Cin, Cout = 2, 3, 4
    x = torch.randn(B, Cin, H, W, dtype=torch.float64, requires_grad=True)
    w = torch.randn(Cout, Cin, KH, KW, dtype=torch.float64)
    y = F.conv2d(x, w, None, 2, (pad, pad_w))
    gy = torch.randn_like(y)
    y.backward(gy)
    Ho, Wo = gy.shape[-2:]
    gx = torch.zeros(B, Cin, H, W, dtype=torch.float64)
    seen = set()
    for py, px, taps in plan(KH, KW, pad):
        assert (py, px) not in seen and taps
        seen.add((py, px))
        Hc, Wc = (H - py + 1) // 2, (W - px + 1) // 2          # pixels of this class
        for dy, dx, tap in taps:
            kh, kw = KH - 1 - tap // KW, KW - 1 - tap % KW       # the dgrad-packed bank is flipped
            shifted = torch.zeros(B, Cout, Hc, Wc, dtype=torch.float64)   # gy[i + dy, j + dx], zero outside (TMA fill)
            i0, i1 = max(0, -dy), min(Hc, Ho - dy)
            j0, j1 = max(0, -dx), min(Wc, Wo - dx)
            if i1 > i0 and j1 > j0:
                shifted[:, :, i0:i1, j0:j1] = gy[:, :, i0 + dy:i1 + dy, j0 + dx:j1 + dx]
            gx[:, :, py::2, px::2] += torch.einsum("bohw,oc->bchw", shifted, w[:, :, kh, kw])
    assert torch.allclose(gx, x.grad, rtol=1e-12, atol=1e-12)


def test_plan_shapes():
    assert [(py, px, len(t)) for py, px, t in plan(3, 3, 1)] == [(0, 0, 1), (0, 1, 2), (1, 0, 2), (1, 1, 4)]
    assert [(py, px, len(t)) for py, px, t in plan(1, 1, 0)] == [(0, 0, 1)]          # three empty classes: caller zero-fills
    assert sum(len(t) for _, _, t in plan(7, 7, 3)) == 49
