import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mono_vifi_b200 import trainer as TR, layers as L, conv
import torch.nn as nn
dev = torch.device("cuda:0")
opt = TR.Options(batch_size=2, height=64, width=128)
torch.manual_seed(0)
models = TR.build_models(opt, dev)
inp = TR.synthetic_inputs(opt, device=dev)

def try_capture(name, fn):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2): fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g, stream=s):
            fn()
        g.replay(); torch.cuda.synchronize()
        print(name, "OK")
    except Exception as e:
        print(name, "FAIL", str(e).splitlines()[0])
        torch.cuda.synchronize()

x = inp[("color_aug", 0, 0)]
def zero():
    for m in models.values():
        for p in m.parameters(): p.grad = None
for backend in ("cudnn", "tcgen05"):
    conv.set_backend(backend)
    enc, dec = models["encoder"], models["depth"]
    def f_conv1():
        zero(); y = enc.encoder.conv1((x - 0.45) / 0.225); y.mean().backward()
    def f_stem():
        zero(); y = enc.encoder.relu(enc.encoder.bn1(enc.encoder.conv1((x - 0.45) / 0.225))); y.mean().backward()
    def f_pool():
        zero(); y = enc.encoder.maxpool(enc.encoder.relu(enc.encoder.bn1(enc.encoder.conv1((x - 0.45) / 0.225)))); y.mean().backward()
    def f_enc():
        zero(); y = enc(x)[-1]; y.mean().backward()
    def f_encdec():
        zero(); y = dec(enc(x))[("disp", 0)]; y.mean().backward()
    def f_pose():
        zero(); p, pi = TR.predict_poses(models, x, x); (p.sum() + pi.sum()).backward()
    for n, f in (("conv1", f_conv1), ("stem", f_stem), ("pool", f_pool), ("enc", f_enc), ("encdec", f_encdec), ("pose", f_pose)):
        try_capture(backend + ":" + n, f)
def f_loss():
    zero()
    out = TR.single_frame_losses(models, inp, opt); out["loss"].backward()
try_capture("full fwd+bwd", f_loss)
