"""Deterministic affine-augmentation inputs for the step-level parity test (shared by tests/golden/gen_step_golden.py and
tests/test_trainstep_cuda.py).  Geometry follows datasets/mono_dataset.py:110-136 (resize ratio, crop origin, angle ->
Rc, box, ratio_local); the images and masks a dataset item would carry are produced with the batched transforms of
mono_vifi_b200/affine.py -- they are INPUTS of the comparison, the expected losses come from the reference."""
import numpy as np
import torch

from mono_vifi_b200 import affine as A

ANGLES = [6.5, -4.0, 2.25, -8.0]
RATIOS = [1.35, 1.8, 1.25, 1.6]
ORIGIN = [(0.3, 0.6), (0.8, 0.1), (0.5, 0.5), (0.0, 1.0)]   # crop origin as a fraction of the slack


def add_affine_inputs(inp, B, H, W):
    angle = torch.tensor([[ANGLES[b % 4]] for b in range(B)])
    ratio = torch.tensor([[RATIOS[b % 4]] for b in range(B)])
    boxes, Rcs = [], []
    for b in range(B):
        r, a = float(ratio[b, 0]), float(angle[b, 0])
        Hre, Wre = int(H * r), int(W * r)
        w0, h0 = int((Wre - W) * ORIGIN[b % 4][0]), int((Hre - H) * ORIGIN[b % 4][1])
        fs = 1 / r
        R = torch.tensor([[np.cos(-np.pi / 180 * a), np.sin(np.pi / 180 * a), 0],
                          [np.sin(-np.pi / 180 * a), np.cos(-np.pi / 180 * a), 0], [0, 0, 1]]).float()
        tmp = R @ torch.tensor([-fs * Wre / 2, -fs * Hre / 2, fs - 1]) + torch.tensor([(Wre / 2 - w0) * fs, (Hre / 2 - h0) * fs, 0])
        K3, iK3 = inp[("K", 0)][b, :3, :3].cpu(), inp[("inv_K", 0)][b, :3, :3].cpu()
        Rc = iK3 @ R @ K3
        Rc[:, 2] += iK3 @ tmp
        Rcs.append(Rc)
        boxes.append([round(w0 / r), round(h0 / r), round(W / r), round(H / r)])
    box = torch.tensor(boxes)
    cpu = {k: v.cpu() for k, v in inp.items()}
    extra = {"Rc": torch.stack(Rcs), "ratio_local": ratio, "angle": angle, "box": box}
    rec = (A.affine_transform(torch.ones(B, 1, H, W), angle, box) > 0).float()
    extra["valid_mask_rec"] = rec
    extra["valid_mask_cons"] = (A.depth_restore(rec, angle, box, torch.ones(B, 1)) > 0).float()
    for f in (-1, 0, 1):
        extra[("color_affine", f, 0)] = A.affine_transform(cpu[("color", f, 0)], angle, box)
    extra[("color_affine_aug", 0, 0)] = A.affine_transform(cpu[("color_aug", 0, 0)], angle, box)
    dev = inp[("K", 0)].device
    inp.update({k: v.to(dev) for k, v in extra.items()})
    return inp
