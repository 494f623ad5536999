"""Golden OUTPUTS and PARAMETER GRADIENTS of the reference's trainable networks in train mode (run in the build
container only: imports the unmodified reference from /root/reference).

For each network: weights filled by name (tests/net_fill.py), seeded inputs, forward in train mode (BatchNorm on batch
statistics, DropPath off through drop_path_rate=0), loss = sum(output * R) with a seeded R, backward through the
reference's autograd.  Stored in tests/golden/netgrad_<name>.npz:
  out_*            outputs (sub-sampled where large)
  gsum, gabs       per parameter (in state_dict order of the parameters): sum and abs-sum of its gradient (float64)
  g_first, g_last  the gradient (sub-sampled above 50k elements) of the first and of the last parameter tensor
  bn_mean, bn_var  running statistics of the first BatchNorm after the step (momentum update of the batch statistics)
Consumed by tests/test_networks_cuda.py (-m gpu): the tcgen05 path has to reproduce them.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import ref_harness  # noqa: E402
import net_fill  # noqa: E402
import netgrad_cases as NC  # noqa: E402

T = ref_harness.import_reference(192, 640, 2)
import networks  # noqa: E402  (the reference's)
import torch  # noqa: E402

torch.set_num_threads(8)


def main():
    for name in NC.CASES:
        torch.manual_seed(0)
        mods, run = NC.build(name, networks)
        for m in mods:
            net_fill.fill_(m, scale=NC.FILL_SCALE.get(name, 1.0))
            m.train()
        outs = run(mods)
        loss = NC.loss_of(outs)
        loss.backward()
        rec = NC.record(mods, outs)
        rec["loss"] = np.float64(loss.item())
        path = os.path.join(HERE, "netgrad_%s.npz" % name)
        np.savez_compressed(path, **rec)
        print("%-18s loss %.6f  %d params  -> %s (%.0f KB)" % (name, rec["loss"], len(rec["gsum"]), os.path.basename(path),
                                                               os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
