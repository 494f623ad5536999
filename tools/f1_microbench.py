"""Times the fused photometric-loss kernels (fwd, bwd) alone with CUDA events; rotates through input sets
larger than L2 so every launch reads from HBM.  Usage: python tools/f1_microbench.py [B H W] [iters]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from mono_vifi_b200 import fused  # noqa: E402
import synth  # noqa: E402


def main():
    smooth = "--smooth" in sys.argv  # low-frequency disparity (what a depth network emits) instead of white noise
    # --coherent: low-frequency disparity without pixel noise and a 10x smaller camera motion (sub-tile displacements,
    # neighbouring pixels sample neighbouring source pixels): the regime of a real depth network / small baseline
    coherent = "--coherent" in sys.argv
    smooth = smooth or coherent
    argv = [a for a in sys.argv if not a.startswith("--")]
    B, H, W = (int(x) for x in argv[1:4]) if len(argv) >= 4 else (12, 192, 640)
    iters = int(argv[4]) if len(argv) >= 5 else 50
    dev = torch.device("cuda:0")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    px = B * H * W
    nset = max(3, int(2.5 * 126e6 / (48 * px)) + 1)
    g = torch.Generator(device=dev).manual_seed(1234)
    K, inv_K = synth.kitti_K(B, H, W)
    c = synth.make_case(1, B, 8, 8, structured=False)
    import tests_helpers  # noqa
    sets = []
    for s in range(nset):
        disp = torch.rand(B, 1, H, W, device=dev, generator=g)
        if smooth:
            lo = torch.rand(B, 1, H // 16, W // 16, device=dev, generator=g)
            disp = (0.0 if coherent else 0.05) * disp + 0.9 * torch.nn.functional.interpolate(lo, size=(H, W), mode="bilinear", align_corners=False) + 0.02
        imgs = [torch.rand(B, 3, H, W, device=dev, generator=g) for _ in range(3)]
        noise = torch.randn(B, 2, H, W, device=dev, generator=g)   # the tie-break noise of train.py:1023 (8 B/px)
        sets.append((disp, *imgs, noise))
    ts = 0.1 if coherent else 1.0
    T = [tests_helpers.synth_T(c["axisangle"][k] * ts, c["translation"][k] * ts, k == 1) for k in range(2)]
    P = [torch.from_numpy(np.matmul(K, T[k])[:, :3, :].astype(np.float32)).to(dev) for k in range(2)]
    invK = torch.from_numpy(inv_K).to(dev)

    def fwd(s):
        d, t, s0, s1, nz = sets[s % nset]
        return fused.f1_forward_raw(d, t, s0, s1, invK, P[0], P[1], nz)

    outs = [fwd(s) for s in range(nset)]
    for s in range(3):
        fused.f1_backward_raw(outs[s]["_saved"], outs[s]["idx"], outs[s]["stats"])
    torch.cuda.synchronize()
    res = {}

    # the launches are recorded into a CUDA graph and replayed, so that the host cost of a call (allocations, six
    # tensor-map encodings, ctypes: ~30-40 us) does not hide kernels of that length
    def timed(fn):
        st = torch.cuda.Stream()
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            fn(0)
        torch.cuda.current_stream().wait_stream(st)
        torch.cuda.synchronize()
        g_ = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_, stream=st):
            for i in range(iters):
                fn(i)
        g_.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g_.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / iters

    res["fwd_us"] = timed(lambda i: fwd(i))

    def bwd(i):
        o = outs[i % nset]
        fused.f1_backward_raw(o["_saved"], o["idx"], o["stats"])
    res["bwd_us"] = timed(bwd)
    res["fwd_GBs"] = 48.0 * px / res["fwd_us"] / 1e3
    res["bwd_GBs"] = 44.0 * px / res["bwd_us"] / 1e3
    res["fwd_frac"] = res["fwd_GBs"] / hbm
    res["bwd_frac"] = res["bwd_GBs"] / hbm
    res.update(lib=os.environ.get("MVF_LIB", "default").split("/")[-1], disp="coherent" if coherent else ("smooth" if smooth else "white-noise"), B=B, H=H, W=W, nset=nset, iters=iters, hbm_peak=hbm, loss=float(outs[0]["loss"][0]))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
