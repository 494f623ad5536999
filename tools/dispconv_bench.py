"""Times the disparity-head kernels (csrc/dispconv.cu) at B12 192x640, C = 16, graph-replayed; python tools/dispconv_bench.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mono_vifi_b200 import _lib
B, C, H, W = 12, 16, 192, 640
L = _lib.lib()
dev = "cuda"
nbuf = 4
xps = [torch.randn(B, H + 2, W + 2, C, device=dev) for _ in range(nbuf)]
gxs = [torch.empty(B, H + 2, W + 2, C, device=dev) for _ in range(nbuf)]
gys = [torch.randn(B, H, W, device=dev) for _ in range(nbuf)]
ys = [torch.empty(B, H, W, device=dev) for _ in range(nbuf)]
w = torch.randn(1, C, 3, 3, device=dev)
b = torch.randn(1, device=dev)
gw = torch.empty(1, C, 3, 3, device=dev)
gb = torch.empty(1, device=dev)
ws = torch.empty(L.mvf_dispconv_wgrad_workspace_floats(B * H * W, C), device=dev)


def timed(run, iters=10):
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        run(0, st.cuda_stream)
    torch.cuda.current_stream().wait_stream(st)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st):
        for i in range(iters):
            run(i, st.cuda_stream)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


f = timed(lambda i, s: _lib.check(L.mvf_dispconv_fwd(xps[i % nbuf].data_ptr(), w.data_ptr(), b.data_ptr(), ys[i % nbuf].data_ptr(), B, C, H, W, s), "fwd"))
d = timed(lambda i, s: _lib.check(L.mvf_dispconv_dgrad(gys[i % nbuf].data_ptr(), w.data_ptr(), gxs[i % nbuf].data_ptr(), B, C, H, W, s), "dgrad"))
g = timed(lambda i, s: _lib.check(L.mvf_dispconv_wgrad(xps[i % nbuf].data_ptr(), gys[i % nbuf].data_ptr(), gw.data_ptr(), gb.data_ptr(), ws.data_ptr(), ws.numel(), B, C, H, W, s), "wgrad"))
mb = B * (H + 2) * (W + 2) * C * 4 / 1e6
print("dispconv B12 16ch 192x640: fwd %.1f us (%.0f GB/s)  dgrad %.1f us (%.0f GB/s)  wgrad %.1f us (%.0f GB/s)" % (f, mb / f * 1e3, d, mb / d * 1e3, g, mb / g * 1e3))
