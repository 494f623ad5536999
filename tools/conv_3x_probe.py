"""3xTF32 accuracy of one convolution (forward, data and weight gradient) against fp64: python tools/conv_3x_probe.py B Cin H W Cout k pad stride"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mono_vifi_b200 import conv_tc
B, Cin, H, W, Cout, k, pad, stride = (int(v) for v in sys.argv[1:9])
g = torch.Generator(device="cuda").manual_seed(9)
x = torch.randn(B, H, W, Cin, device="cuda", generator=g).permute(0, 3, 1, 2).requires_grad_(True)
w = (torch.randn(Cout, Cin, k, k, device="cuda", generator=g) / (Cin * k * k) ** 0.5).requires_grad_(True)
with conv_tc.precision("3xtf32"):
    y = conv_tc.conv2d(x, w, None, stride, pad)
    gy = torch.randn(y.shape, device="cuda", generator=g)
    y.backward(gy)
torch.cuda.synchronize()
xr, wr = x.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True)
yr = torch.nn.functional.conv2d(xr, wr, None, stride, pad)
yr.backward(gy.double())
for name, got, ref in (("y", y, yr), ("gx", x.grad, xr.grad), ("gw", w.grad, wr.grad)):
    print("%s rel err %.3g" % (name, (got.double() - ref).abs().max().item() / ref.abs().max().item()), end="   ")
print()
