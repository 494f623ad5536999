import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mono_vifi_b200 import trainer as TR
dev = torch.device("cuda:0")
mode = sys.argv[1]
opt = TR.Options(batch_size=2, height=64, width=128)
torch.manual_seed(0)
step = TR.TrainStep(opt, dev, capturable=True)
step.train()
inp = TR.synthetic_inputs(opt, device=dev)
if "eagerfirst" in mode:
    step(inp); torch.cuda.synchronize()
def fn():
    if "fb" in mode:
        step.optimizer.zero_grad(set_to_none=True)
        out = TR.single_frame_losses(step.models, inp, opt); out["loss"].backward()
    elif "clip" in mode:
        step.forward_backward(inp)
        torch.nn.utils.clip_grad_norm_([p for p in step.params if p.grad is not None], 5.0)
    else:
        step(inp)
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2): fn()
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g, stream=s):
        fn()
    g.replay(); torch.cuda.synchronize(); print(mode, "OK")
except Exception as e:
    print(mode, "FAIL", str(e).splitlines()[0])
