"""One convolution problem through the tcgen05 kernel vs fp64: python tools/conv_case.py B Cin H W Cout k pad stride"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mono_vifi_b200 import conv_tc
B, Cin, H, W, Cout, k, pad, stride = (int(a) for a in sys.argv[1:9])
g = torch.Generator(device="cuda").manual_seed(3)
x = torch.randn(B, Cin, H, W, device="cuda", generator=g)
w = torch.randn(Cout, Cin, k, k, device="cuda", generator=g) / (Cin * k * k) ** 0.5
try:
    y = conv_tc.conv_forward_raw(x, conv_tc.pack_filters(w), None, Cout, k, k, pad, stride)
    torch.cuda.synchronize()
except Exception as e:
    print(sys.argv[1:], "EXC", str(e).splitlines()[0]); sys.exit(0)
ref = torch.nn.functional.conv2d(x.double(), w.double(), None, stride, pad).float()
print(sys.argv[1:], "max err %.4g of %.4g" % ((y - ref).abs().max().item(), ref.abs().max().item()))
