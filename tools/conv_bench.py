"""Times the tcgen05 convolution kernels (fprop / wgrad) on the ResNet18 layer shapes of the benchmark (B12 192x640)
with CUDA events; inputs rotate over buffers larger than L2.  python tools/conv_bench.py [iters]"""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mono_vifi_b200 import conv_tc
SHAPES = [  # name, B, Cin, H, W, Cout, k, pad, stride
    ("layer1 64->64 48x160", 12, 64, 48, 160, 64, 3, 1, 1),
    ("layer2 128->128 24x80", 12, 128, 24, 80, 128, 3, 1, 1),
    ("layer3 256->256 12x40", 12, 256, 12, 40, 256, 3, 1, 1),
    ("layer4 512->512 6x20", 12, 512, 6, 20, 512, 3, 1, 1),
    ("up1,1 96->32 96x320 (valid)", 12, 96, 98, 322, 32, 3, 0, 1),
    ("up0,1 16->16 192x640 (valid)", 12, 16, 194, 642, 16, 3, 0, 1),
    ("layer2.0 64->128 s2", 12, 64, 48, 160, 128, 3, 1, 2),
    ("stem 4->64 7x7 s2", 12, 4, 192, 640, 64, 7, 3, 2),
]
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
only = sys.argv[2] if len(sys.argv) > 2 else None
dev = "cuda"
for name, B, Cin, H, W, Cout, k, pad, stride in SHAPES:
    if only and only not in name:
        continue
    nbuf = max(2, int(300e6 / (B * Cin * H * W * 4)) + 1)
    nbuf = min(nbuf, 8)
    xs = [torch.randn(B, H, W, Cin, device=dev).permute(0, 3, 1, 2) for _ in range(nbuf)]
    w = torch.randn(Cout, Cin, k, k, device=dev) / (Cin * k * k) ** 0.5
    wp = conv_tc.pack_filters(w)
    Ho, Wo = conv_tc.out_hw(H, W, k, k, pad, stride)
    gys = [torch.randn(B, Ho, Wo, Cout, device=dev).permute(0, 3, 1, 2) for _ in range(nbuf)]
    flops = 2.0 * B * Ho * Wo * Cout * Cin * k * k
    res = {}
    for what in os.environ.get("CONV_BENCH_WHAT", "fprop,wgrad").split(","):
        def run(i):
            if what == "fprop":
                conv_tc.conv_forward_raw(xs[i % nbuf], wp, None, Cout, k, k, pad, stride)
            else:
                conv_tc.weight_grad(xs[i % nbuf], gys[i % nbuf], (Cout, Cin, k, k), pad, stride)
        # the calls are recorded into a CUDA graph so that the host-side cost of a call (tensor-map encoding, ctypes,
        # allocator) does not hide kernels shorter than ~40 us
        st = torch.cuda.Stream()
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            for i in range(3):
                run(i)
        torch.cuda.current_stream().wait_stream(st)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for i in range(iters):
                run(i)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / iters
        res[what] = (us, flops / us / 1e6)
    print("%-32s %s | %.1f GF" % (name, " | ".join("%s %7.1f us %6.1f TF/s" % (k, v[0], v[1]) for k, v in res.items()), flops / 1e9))
