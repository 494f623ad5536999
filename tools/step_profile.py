"""Kernel-time breakdown of one eager training step (torch.profiler / CUPTI): python tools/step_profile.py <config> [top N]
Groups kernels by name; prints total time, launches and share.  (A profiler run: shares, not bench values.)"""
import collections
import os
import re
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mono_vifi_b200 import trainer as TR  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "mf"
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
dev = torch.device("cuda:0")
opt = TR.Options(**bench.CONFIGS[cfg]["opt"])
torch.manual_seed(1234)
step = TR.TrainStep(opt, dev)
step.train()
inputs = TR.synthetic_inputs(opt, dev)
for _ in range(3):
    step(inputs)
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    step(inputs)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0.0, 0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        name = re.sub(r"<.*", "", e.name)
        name = re.sub(r"^void ", "", name)
        a = agg[name[:70]]
        a[0] += e.device_time
        a[1] += 1
tot = sum(a[0] for a in agg.values())
print("config %s: %.2f ms of kernel time in %d launches" % (cfg, tot / 1e3, sum(a[1] for a in agg.values())))
for name, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%6.2f%% %9.1f us %6d  %s" % (100 * t / tot, t, n, name))
