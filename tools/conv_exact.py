"""Exactness probe: small-integer inputs make every TF32 product and every fp32 partial sum exact, so the tensor-core convolutions must
equal torch's fp64 result BIT FOR BIT whatever the tile shape, issuer split or split-K order (any difference is a real bug, not
rounding).  python tools/conv_exact.py            (prints one line per case; exit code 1 on a mismatch)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mono_vifi_b200 import conv_tc
CASES = [  # B, Cin, H, W, Cout, k, pad, stride
    (6, 256, 4, 6, 512, 3, 1, 2), (12, 256, 4, 6, 512, 3, 1, 2), (6, 256, 4, 6, 512, 1, 0, 2), (12, 256, 4, 6, 512, 1, 0, 2),
    (6, 512, 2, 3, 512, 3, 1, 1), (12, 512, 2, 3, 512, 3, 1, 1), (2, 64, 16, 24, 128, 3, 1, 2), (2, 64, 16, 24, 128, 1, 0, 2),
    (6, 64, 16, 24, 64, 3, 1, 1), (12, 128, 8, 12, 128, 3, 1, 1), (3, 512, 2, 3, 256, 1, 0, 1), (2, 8, 64, 96, 64, 7, 3, 2),
    (12, 64, 48, 160, 128, 3, 1, 2), (4, 16, 32, 48, 16, 3, 0, 1), (2, 96, 18, 26, 32, 3, 0, 1),
]
bad = 0
g = torch.Generator(device="cuda").manual_seed(3)
for B, Cin, H, W, Cout, k, pad, stride in CASES:
    x = torch.randint(-2, 3, (B, H, W, Cin), device="cuda", generator=g).float().permute(0, 3, 1, 2)
    w = torch.randint(-2, 3, (Cout, Cin, k, k), device="cuda", generator=g).float()
    Ho, Wo = conv_tc.out_hw(H, W, k, k, pad, stride)
    gy = torch.randint(-2, 3, (B, Ho, Wo, Cout), device="cuda", generator=g).float().permute(0, 3, 1, 2)
    ref_y = torch.nn.functional.conv2d(x.double(), w.double(), None, stride, pad)
    ref_gx, ref_gw, _ = torch.ops.aten.convolution_backward(gy.double(), x.double(), w.double(), None, [stride, stride], [pad, pad], [1, 1],
                                                            False, [0, 0], 1, [True, True, False])
    res = {}
    for rep in range(3):
        res["fprop"] = (conv_tc.conv_forward_raw(x, conv_tc.pack_filters(w), None, Cout, k, k, pad, stride), ref_y)
        res["wgrad"] = (conv_tc.weight_grad(x, gy, (Cout, Cin, k, k), pad, stride), ref_gw)
        if stride == 1:
            res["dgrad"] = (conv_tc.conv_forward_raw(gy, conv_tc.pack_filters(w, dgrad=True), None, Cin, k, k, k - 1 - pad), ref_gx)
        elif Cin % 4 == 0:
            res["dgrad"] = (conv_tc.input_grad_s2(x, gy, w, pad), ref_gx)
        torch.cuda.synchronize()
        for name, (got, ref) in res.items():
            d = (got.double() - ref).abs().max().item()
            if d != 0.0:
                bad += 1
                print("MISMATCH %s rep %d %s: max |diff| %g (max |ref| %g)" % (name, rep, (B, Cin, H, W, Cout, k, pad, stride), d, ref.abs().max().item()), flush=True)
    print("case", (B, Cin, H, W, Cout, k, pad, stride), "done", flush=True)
print("mismatches:", bad)
sys.exit(1 if bad else 0)
