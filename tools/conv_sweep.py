"""Tile-shape sweep of the stride-1 patch convolution kernel: for each layer shape, every (N_TILE, MT) the kernel is instantiated
for (MVF_CONV_TILE_RULE=N,MT), full kernel / no MMAs (MVF_CONV_DBG=4: the data movement alone) / no TMA loads (8: MMAs + epilogue
alone); graph-replayed launches over rotating buffers.  python tools/conv_sweep.py [iters] [name filter]
Feeds the cost model in csrc/conv_tc.cu (conv_forward_patch)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mono_vifi_b200 import conv_tc
SHAPES = [  # name, B, Cin, H, W, Cout, k, pad
    ("64->64 48x160", 12, 64, 48, 160, 64, 3, 1),
    ("128->128 24x80", 12, 128, 24, 80, 128, 3, 1),
    ("256->256 12x40", 12, 256, 12, 40, 256, 3, 1),
    ("512->512 6x20", 12, 512, 6, 20, 512, 3, 1),
    ("16->16 192x640 valid", 12, 16, 194, 642, 16, 3, 0),
    ("96->32 96x320 valid", 12, 96, 98, 322, 32, 3, 0),
]
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
only = sys.argv[2] if len(sys.argv) > 2 else None
extra_env = {}
dev = "cuda"


def timed(run):
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        for i in range(2):
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
    return e0.elapsed_time(e1) * 1e3 / iters


for name, B, Cin, H, W, Cout, k, pad in SHAPES:
    if only and only not in name:
        continue
    nbuf = min(8, max(2, int(300e6 / (B * Cin * H * W * 4)) + 1))
    xs = [torch.randn(B, H, W, Cin, device=dev).permute(0, 3, 1, 2) for _ in range(nbuf)]
    w = torch.randn(Cout, Cin, k, k, device=dev) / (Cin * k * k) ** 0.5
    wp = conv_tc.pack_filters(w)
    Ho, Wo = conv_tc.out_hw(H, W, k, k, pad, 1)
    flops = 2.0 * B * Ho * Wo * Cout * Cin * k * k
    n_wide = 16
    while n_wide < Cout and n_wide < 128:
        n_wide *= 2
    rules = ["r1", None] + ["%d,%d" % (n, mt) for n in (128, 64, 32, 16) if n <= n_wide for mt in (1, 2)]
    for rule in rules:
        row = []
        for dbg in (0, 4, 8):
            if rule is None:
                os.environ.pop("MVF_CONV_TILE_RULE", None)
            else:
                os.environ["MVF_CONV_TILE_RULE"] = rule
            os.environ["MVF_CONV_DBG"] = str(dbg)
            try:
                us = timed(lambda i: conv_tc.conv_forward_raw(xs[i % nbuf], wp, None, Cout, k, k, pad, 1))
            except Exception as e:  # shape does not fit in shared memory
                us = float("nan")
            row.append(us)
        print("%-22s rule %-6s  full %6.1f us %6.1f TF/s | no-MMA %6.1f us | no-TMA %6.1f us" % (
            name, rule or "model", row[0], flops / row[0] / 1e6, row[1], row[2]), flush=True)
os.environ.pop("MVF_CONV_TILE_RULE", None)
os.environ.pop("MVF_CONV_DBG", None)
