"""Per-shape time of the tensor-core convolution kernels in one eager, single-stream training step (CUDA events around every
launch): python tools/conv_layers.py [config] -- which layers hold the step's convolution time and how far each is from the
TF32 peak.  Stride-1 dgrad launches are listed under the shape of the convolution they evaluate (gy as input)."""
import collections
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mono_vifi_b200 import conv_tc, trainer as TR  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "2"
dev = torch.device("cuda:0")
opt = TR.Options(**bench.CONFIGS[cfg]["opt"])
torch.manual_seed(1234)
step = TR.TrainStep(opt, dev)
step.train()
step.side = step.side2 = None
inputs = TR.synthetic_inputs(opt, dev)
for _ in range(3):
    step(inputs)
conv_tc.timing, conv_tc.shapes = [], []
torch.cuda._sleep(int(4e8))   # ~0.2 s: the host enqueues the whole step behind it, so every event interval is pure GPU time
step(inputs)
torch.cuda.synchronize()
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops_sustained", 1400.0) / 2 if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 700.0
agg = collections.defaultdict(lambda: [0.0, 0.0, 0])
for (tag, fl, a, b), shp in zip(conv_tc.timing, conv_tc.shapes):
    e = agg[(tag, shp)]
    e[0] += a.elapsed_time(b) * 1e3
    e[1] += fl
    e[2] += 1
tot = sum(e[0] for e in agg.values())
print("config %s: %.2f ms in %d tensor-core convolution launches; TF32 peak %.0f TFLOP/s" % (cfg, tot / 1e3, sum(e[2] for e in agg.values()), peak))
print("%-6s %-44s %5s %9s %8s %7s %6s" % ("kind", "(B, Cin, H, W, Cout, KH, KW, stride)", "n", "us total", "us each", "TF/s", "share"))
for (tag, shp), (us, fl, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%-6s %-44s %5d %9.1f %8.1f %7.1f %5.1f%%" % (tag, str(shp), n, us, us / n, fl / us / 1e6, 100 * us / tot))
