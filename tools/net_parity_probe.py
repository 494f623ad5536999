"""Measures, on the GPU, how far each arithmetic class lands from the reference's CPU fp32 results on the
train-mode network cases of tests/netgrad_cases.py (fixtures: tests/golden/netgrad_*.npz):

  3xtf32       tcgen05 kernels, three tensor-core products per product (fp32 class)           -- ours
  tf32         tcgen05 kernels, single TF32 products (production)                              -- ours
  lib-fp32     cuDNN with allow_tf32 = False + torch BatchNorm: the GPU fp32 noise floor        -- library
  lib-tf32     cuDNN with allow_tf32 = True (what the reference runs by default on a GPU)      -- library

Prints, per case and class: max output error / output scale, max per-tensor |sum(grad) - sum(ref)| / abs-sum(ref),
max per-tensor abs-sum deviation, element-wise error of the first / last gradient tensor.  The tolerances of
tests/test_networks_cuda.py are set from this table (profiles/r2_net_parity.md).
python tools/net_parity_probe.py [case ...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import net_fill  # noqa: E402
import netgrad_cases as NC  # noqa: E402
from mono_vifi_b200 import bn_act, conv, conv_tc, networks  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def run(name, cls):
    torch.manual_seed(0)
    mods, fn = NC.build(name, networks)
    for m in mods:
        net_fill.fill_(m)
        m.cuda().train()
    old_backend, old_bn = conv.get_backend(), bn_act.enabled
    try:
        if cls.startswith("lib"):
            conv.set_backend("cudnn")
            bn_act.enabled = False
            torch.backends.cudnn.allow_tf32 = cls == "lib-tf32"
            outs = fn(mods, "cuda")
            NC.loss_of(outs).backward()
        else:
            with conv_tc.precision(cls):
                outs = fn(mods, "cuda")
                NC.loss_of(outs).backward()
    finally:
        conv.set_backend(old_backend)
        bn_act.enabled = old_bn
    torch.cuda.synchronize()
    return NC.record(mods, outs), len(outs)


def main():
    cases = sys.argv[1:] or NC.CASES
    print("| case | class | out err/scale | grad sum err/abs | grad abs-sum dev | g_first | g_last | worst tensor |")
    print("|---|---|---|---|---|---|---|---|")
    for name in cases:
        g = np.load(os.path.join(GOLD, "netgrad_%s.npz" % name))
        for cls in ("3xtf32", "tf32", "lib-fp32", "lib-tf32"):
            rec, n = run(name, cls)
            eo = max(float(np.abs(rec["out_%d" % i] - g["out_%d" % i]).max()) / max(1e-6, float(np.abs(g["out_%d" % i]).max()))
                     for i in range(n))
            den = np.maximum(g["gabs"], 1e-12)
            es = np.abs(rec["gsum"] - g["gsum"]) / den
            ea = np.abs(rec["gabs"] - g["gabs"]) / den
            es[g["gabs"] == 0] = 0
            ea[g["gabs"] == 0] = 0
            w = int(np.argmax(np.maximum(es, ea)))
            ef = float(np.abs(rec["g_first"] - g["g_first"]).max()) / max(1e-9, float(np.abs(g["g_first"]).max()))
            el = float(np.abs(rec["g_last"] - g["g_last"]).max()) / max(1e-9, float(np.abs(g["g_last"]).max()))
            print("| %s | %s | %.1e | %.1e | %.1e | %.1e | %.1e | %s |" % (name, cls, eo, es.max(), ea.max(), ef, el, g["names"][w]))
        sys.stdout.flush()


if __name__ == "__main__":
    main()
