"""Structured probes of the tcgen05 convolution kernel (prints what the kernel did with recognisable inputs)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mono_vifi_b200 import conv_tc  # noqa: E402


def run(x, w, pad=0):
    Cout, Cin, k, _ = w.shape
    y = conv_tc.conv_forward_raw(x, conv_tc.pack_filters(w), None, Cout, k, k, pad)
    torch.cuda.synchronize()
    return y


def main():
    torch.set_printoptions(linewidth=200, precision=3, sci_mode=False)
    dev = "cuda"
    import ctypes
    from mono_vifi_b200 import _lib
    dbg = torch.full((4096 + 128 * 32 + 8,), -7.0, device=dev)
    L = ctypes.CDLL(_lib.SO_PATH)
    L.mvf_conv2d_debug_buffer(ctypes.c_void_p(dbg.data_ptr()))
    B, Cin, H, W, Cout = 1, 8, 4, 32, 16
    x = torch.zeros(B, Cin, H, W, device=dev)
    for c in range(Cin):
        x[0, c] = (c + 1) * 1000 + torch.arange(W, device=dev).float()[None, :] + 100 * torch.arange(H, device=dev).float()[:, None]
    w = torch.zeros(Cout, Cin, 1, 1, device=dev)
    for n in range(Cout):
        w[n, :, 0, 0] = (n + 1) * 10 + torch.arange(Cin, device=dev).float()
    y = run(x, w)
    d = dbg.cpu()
    print("T0 A stage dump, first 1024 B atom (8 rows of 32):")
    print(d[:256].view(8, 32)[:, :12])
    print("A dump atom 1 (tile row 1):")
    print(d[256:512].view(8, 32)[:, :12])
    print("A dump k-group 1 (expect zeros):", d[1024:1032].tolist())
    print("B stage dump (16 rows of 32 k):")
    print(d[4096:4096 + 512].view(16, 32)[:, :12])
    print("tmem base: 0x%08x" % d[4096 + 512].view(torch.int32).item())
    print("y[0,:4,0,:6]:", y[0, :4, 0, :6])
    x = torch.ones(B, Cin, H, W, device=dev)
    w = torch.ones(Cout, Cin, 1, 1, device=dev)
    y = run(x, w)
    print("T1 ones: expect 8 everywhere; got min %.3f max %.3f mean %.3f" % (y.min(), y.max(), y.mean()))
    print(y[0, 0, :, :8])
    x = torch.arange(1, Cin + 1, device=dev).float().view(1, Cin, 1, 1).expand(B, Cin, H, W).contiguous()
    y = run(x, w)
    print("T2 per-channel constants: expect 36; got min %.3f max %.3f" % (y.min(), y.max()))
    x = torch.zeros(B, Cin, H, W, device=dev)
    x[0, 0] = torch.arange(W, device=dev).float()[None, :] + 100 * torch.arange(H, device=dev).float()[:, None]
    w = torch.zeros(Cout, Cin, 1, 1, device=dev)
    w[:, 0] = 1
    y = run(x, w)
    print("T3 pixel pattern (expect w + 100 h):")
    print(y[0, 0, :, :12])
    print(y[0, 5, :, 20:32])
    x = torch.zeros(B, Cin, H, W, device=dev)
    x[0, 0] = 1
    w = torch.zeros(Cout, Cin, 1, 1, device=dev)
    w[:, 0, 0, 0] = torch.arange(1, Cout + 1, device=dev).float()
    y = run(x, w)
    print("T4 cout pattern (expect n+1):", y[0, :, 0, 0].tolist(), y[0, :, 3, 31].tolist())
    x = torch.zeros(B, Cin, H, W, device=dev)
    w = torch.zeros(Cout, Cin, 1, 1, device=dev)
    for c in range(Cin):
        x[0, c] = 1
        w[:, c, 0, 0] = 10 ** c if c < 7 else 0.5
    y = run(x, w)
    print("T5 channel pattern (expect 1111111.5):", y[0, 0, 0, 0].item(), y[0, 3, 2, 7].item())
    # 3x3
    x = torch.randn(1, 32, 8, 64, device=dev)
    w = torch.randn(16, 32, 3, 3, device=dev) / 17.0
    y = run(x, w, 1)
    ref = torch.nn.functional.conv2d(x.double(), w.double(), None, 1, 1).float()
    print("T6 3x3 random: max err %.4g of %.4g" % ((y - ref).abs().max(), ref.abs().max()))


if __name__ == "__main__":
    main()
