"""Debug aid: capture the multi-frame step of a backbone into a CUDA graph at a small size and print the full traceback of a failure.
   python tools/graph_probe_mf.py [ResNet18|DHRNet|LiteMono] [vfi_scale]"""
import os
import sys
import traceback

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mono_vifi_b200 import trainer as TR  # noqa: E402

backbone = sys.argv[1] if len(sys.argv) > 1 else "ResNet18"
vfi = sys.argv[2] if len(sys.argv) > 2 else "small"
dev = torch.device("cuda:0")
opt = TR.Options(batch_size=2, height=64, width=96, multi_frame=True, backbone=backbone, vfi_scale=vfi)
torch.manual_seed(0)
step = TR.TrainStep(opt, dev, capturable=True)
step.train()
inputs = TR.synthetic_inputs(opt, dev)
print("eager", float(step(inputs)))
try:
    g = TR.GraphedTrainStep(step, inputs, warmup=2)
    for _ in range(3):
        l = g(inputs)
    torch.cuda.synchronize()
    print("graph", float(l))
except Exception:
    traceback.print_exc()
