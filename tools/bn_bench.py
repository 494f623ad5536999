"""fused BN(+add)+ReLU vs cuDNN BN + torch add/relu, fwd+bwd, CUDA-graph timed"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mono_vifi_b200 import bn_act
def timeit(fn, iters=20):
    st = torch.cuda.Stream(); st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        for _ in range(3): fn()
    torch.cuda.current_stream().wait_stream(st); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st):
        for _ in range(iters): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters
for (B, C, H, W) in [(12, 64, 96, 320), (12, 64, 48, 160), (12, 128, 24, 80), (12, 256, 12, 40), (12, 512, 6, 20)]:
    x = torch.randn(B, H, W, C, device="cuda").permute(0, 3, 1, 2).requires_grad_(True)
    idn = torch.randn(B, H, W, C, device="cuda").permute(0, 3, 1, 2).requires_grad_(True)
    gy = torch.randn(B, H, W, C, device="cuda").permute(0, 3, 1, 2)
    bn = torch.nn.BatchNorm2d(C).cuda().train()
    def run(fused):
        bn_act.enabled = fused
        x.grad = idn.grad = None
        y = bn_act.bn_act(bn, x, idn, True)
        y.backward(gy)
    t1 = timeit(lambda: run(True)); t0 = timeit(lambda: run(False))
    mb = B * C * H * W * 4 / 1e6
    print("%-20s %.1f MB  fused %.1f us   torch %.1f us" % ((B, C, H, W), mb, t1, t0))
