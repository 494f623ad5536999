"""Which ATen operators still launch kernels inside a training step, and from where: python tools/aten_ops.py <config> [top N]
One eager step under torch.profiler (CPU + CUDA, with stacks); for every ATen op that launched device kernels: calls, device time and
the innermost repo source line that called it.  (A profiler run: shares and counts, not bench values.)"""
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mono_vifi_b200 import trainer as TR  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "2"
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
dev = torch.device("cuda:0")
opt = TR.Options(**bench.CONFIGS[cfg]["opt"])
torch.manual_seed(1234)
step = TR.TrainStep(opt, dev)
step.train()
inputs = TR.synthetic_inputs(opt, dev)
for _ in range(3):
    step(inputs)
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CPU, torch.profiler.ProfilerActivity.CUDA], with_stack=True) as prof:
    step(inputs)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type != torch.autograd.DeviceType.CPU or not e.name.startswith("aten::"):
        continue
    kern = [k for k in e.kernels]
    if not kern:
        continue
    # only leaf ops (an op whose child op launched the kernel would be counted twice)
    if any(c.name.startswith("aten::") and c.kernels for c in e.cpu_children):
        continue
    where = "?"
    for fr in (e.stack or []):
        if "/mono_vifi_b200/" in fr or "/bench.py" in fr:
            where = fr.split("/root/repo/")[-1].split("repo/")[-1][:90]
            break
    a = agg[(e.name, where)]
    a[0] += 1
    a[1] += sum(k.duration for k in kern)
tot_n = sum(a[0] for a in agg.values())
tot_t = sum(a[1] for a in agg.values())
print("config %s: %d ATen ops launched kernels, %.1f us of device time" % (cfg, tot_n, tot_t))
for (name, where), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5d x %-28s %8.1f us  %s" % (n, name, t, where))
