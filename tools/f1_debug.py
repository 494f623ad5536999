"""Debug helper (GPU box): per-pixel comparison of the CUDA backward with the oracle on a golden case."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth  # noqa: E402
from oracle import f1_oracle as O  # noqa: E402
from mono_vifi_b200 import fused  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "small"
use_mask = int(sys.argv[2]) if len(sys.argv) > 2 else 1
g = np.load(os.path.join(ROOT, "tests", "golden", "f1_%s.npz" % name))
seed, B, H, W, structured, with_mask, f0, f1, f2 = [int(x) for x in g["spec"]]
flags = f0 * 1 + f1 * 2 + f2 * 4
c = synth.make_case(seed, B, H, W, bool(structured), bool(with_mask), 1 if f1 else 2)
if not use_mask:
    c["mask_rec"] = None
dev = torch.device("cuda:0")
t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)
out = fused.f1_forward_raw(t(c["disp"]), t(c["tgt"]), t(c["src0"]), t(c["src1"]), t(c["inv_K"]), t(g["P"][0]),
                           t(g["P"][1]), t(c["noise"]), t(c["mask_rec"]), flags=flags)
gd, gP0, gP1 = fused.f1_backward_raw(out["_saved"], out["idx"], out["stats"], None, flags=flags)
torch.cuda.synchronize()
idx = out["idx"].cpu().numpy()
rd, rP0, rP1 = O.f1_backward(c["disp"], c["tgt"], c["src0"], c["src1"], c["inv_K"], g["P"][0], g["P"][1], idx,
                             c["mask_rec"], 1.0, flags=flags)
gd = gd.cpu().numpy()
err = np.abs(gd - rd)[:, 0]
scale = np.abs(rd).max()
print("case", name, "mask", use_mask, "g_disp scale", scale, "max err", err.max(), "mean err", err.mean())
bad = np.argwhere(err > 1e-3 * scale)
print("n bad", len(bad), "of", err.size)
for bb, y, x in bad[:40]:
    print("  b%d y%d x%d (ty %d tx %d) cuda % .4e ref % .4e idx %d" % (bb, y, x, y % 16, x % 32, gd[bb, 0, y, x], rd[bb, 0, y, x], idx[bb, y, x]))
for nm, a, r in (("gP0", gP0.cpu().numpy(), rP0), ("gP1", gP1.cpu().numpy(), rP1)):
    print(nm, "rel err per element:\n", np.array2string((a - r) / (np.abs(r) + 1e-12), precision=2, max_line_width=150))
