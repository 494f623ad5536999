"""Timeline analysis of ONE graph-replayed training step (torch.profiler / CUPTI kernel intervals): how long the GPU runs 0 / 1 / 2+
kernels at once, and, per kernel name, its total time and its EXCLUSIVE time (intervals where it is the only kernel running -- the
part of the step's wall time that only making that kernel faster, or overlapping it, can remove).
python tools/step_timeline.py [config] [top N]      (a profiler run: shares, not bench values)"""
import collections
import os
import re
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mono_vifi_b200 import trainer as TR  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "2"
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
dev = torch.device("cuda:0")
opt = TR.Options(**bench.CONFIGS[cfg]["opt"])
torch.manual_seed(1234)
step = TR.TrainStep(opt, dev)
step.train()
inputs = TR.synthetic_inputs(opt, dev)
g = TR.GraphedTrainStep(step, inputs)
for _ in range(3):
    g()
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    g()
    torch.cuda.synchronize()
iv = []
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA and e.device_time > 0:
        name = re.sub(r"^void ", "", re.sub(r"<.*", "", e.name))
        name = re.sub(r"\(anonymous namespace\)::", "", name)[:60]
        t0 = e.time_range.start
        iv.append((t0, t0 + e.device_time, name))
iv.sort()
t_begin, t_end = iv[0][0], max(b for _, b, _ in iv)
# sweep line
pts = []
for i, (a, b, _) in enumerate(iv):
    pts.append((a, 1, i))
    pts.append((b, -1, i))
pts.sort(key=lambda p: (p[0], p[1]))
active = set()
level_time = collections.Counter()
excl = collections.Counter()
total = collections.Counter()
count = collections.Counter()
for a, b, n in iv:
    total[n] += b - a
    count[n] += 1
prev = pts[0][0]
for t, d, i in pts:
    if t > prev:
        k = len(active)
        level_time[min(k, 4)] += t - prev
        if k == 1:
            excl[iv[next(iter(active))][2]] += t - prev
        prev = t
    if d > 0:
        active.add(i)
    else:
        active.discard(i)
wall = t_end - t_begin
print("config %s: one graph replay = %.2f ms wall (first kernel start to last kernel end), %d kernels, %.2f ms of kernel time" % (
    cfg, wall / 1e3, len(iv), sum(total.values()) / 1e3))
for k in sorted(level_time):
    print("  %s kernels running: %7.2f ms  %5.1f%%" % (("%d" % k) if k < 4 else "4+", level_time[k] / 1e3, 100 * level_time[k] / wall))
print("%-62s %6s %9s %9s %7s" % ("kernel", "n", "total us", "excl us", "excl %"))
for n, t in sorted(excl.items(), key=lambda kv: -kv[1])[:top]:
    print("%-62s %6d %9.1f %9.1f %6.1f%%" % (n, count[n], total[n], t, 100 * t / wall))
