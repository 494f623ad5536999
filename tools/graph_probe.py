"""Eager vs CUDA-graph replay of the training step (1 GPU): step time and loss agreement."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mono_vifi_b200 import trainer as TR
dev = torch.device("cuda:0")
opt = TR.Options()
torch.manual_seed(1234)
step = TR.TrainStep(opt, dev, capturable=True)
step.train()
inp = [TR.synthetic_inputs(opt, device=dev, seed=1234 + s) for s in range(2)]
for i in range(3):
    l = step(inp[i % 2])
torch.cuda.synchronize()
def timeit(fn, n=10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("eager ms/step", timeit(lambda i: step(inp[i % 2])), "loss", float(step(inp[0])))
g = TR.GraphedTrainStep(step, inp[0])
print("graph ms/step", timeit(lambda i: g(inp[i % 2])), "loss", float(g(inp[0])))
print("graph ms/step (no input copy)", timeit(lambda i: g()))
