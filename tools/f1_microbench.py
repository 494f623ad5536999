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
    nset = max(3, int(2.5 * 126e6 / (40 * px)) + 1)
    g = torch.Generator(device=dev).manual_seed(1234)
    K, inv_K = synth.kitti_K(B, H, W)
    c = synth.make_case(1, B, 8, 8, structured=False)
    import tests_helpers  # noqa
    sets = []
    for s in range(nset):
        disp = torch.rand(B, 1, H, W, device=dev, generator=g)
        if smooth:
            lo = torch.rand(B, 1, H // 16, W // 16, device=dev, generator=g)
            disp = 0.05 * disp + 0.9 * torch.nn.functional.interpolate(lo, size=(H, W), mode="bilinear", align_corners=False)
        imgs = [torch.rand(B, 3, H, W, device=dev, generator=g) for _ in range(3)]
        sets.append((disp, *imgs))
    T = [tests_helpers.synth_T(c["axisangle"][k], c["translation"][k], k == 1) for k in range(2)]
    P = [torch.from_numpy(np.matmul(K, T[k])[:, :3, :].astype(np.float32)).to(dev) for k in range(2)]
    invK = torch.from_numpy(inv_K).to(dev)

    def fwd(s):
        d, t, s0, s1 = sets[s % nset]
        return fused.f1_forward_raw(d, t, s0, s1, invK, P[0], P[1])

    outs = [fwd(s) for s in range(nset)]
    for s in range(3):
        fused.f1_backward_raw(outs[s]["_saved"], outs[s]["idx"], outs[s]["stats"])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    res = {}
    e0.record()
    for i in range(iters):
        fwd(i)
    e1.record()
    torch.cuda.synchronize()
    res["fwd_us"] = e0.elapsed_time(e1) * 1e3 / iters
    e0.record()
    for i in range(iters):
        o = outs[i % nset]
        fused.f1_backward_raw(o["_saved"], o["idx"], o["stats"])
    e1.record()
    torch.cuda.synchronize()
    res["bwd_us"] = e0.elapsed_time(e1) * 1e3 / iters
    res["fwd_GBs"] = 40.0 * px / res["fwd_us"] / 1e3
    res["bwd_GBs"] = 44.0 * px / res["bwd_us"] / 1e3
    res["fwd_frac"] = res["fwd_GBs"] / hbm
    res["bwd_frac"] = res["bwd_GBs"] / hbm
    res.update(lib=os.environ.get("MVF_LIB", "default").split("/")[-1], disp="smooth" if smooth else "white-noise", B=B, H=H, W=W, nset=nset, iters=iters, hbm_peak=hbm, loss=float(outs[0]["loss"][0]))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
