"""Deterministic affine-augmentation inputs for the step-level parity test (shared by tests/golden/gen_step_golden.py and
tests/test_trainstep_cuda.py).  Geometry follows datasets/mono_dataset.py:110-136 (resize ratio, crop origin, angle ->
Rc, box, ratio_local); the images and masks a dataset item would carry are produced with the batched transforms of
mono_vifi_b200/affine.py -- they are INPUTS of the comparison, the expected losses come from the reference."""
import torch

from mono_vifi_b200 import affine as A

ANGLES = [6.5, -4.0, 2.25, -8.0]
RATIOS = [1.35, 1.8, 1.25, 1.6]
ORIGIN = [(0.3, 0.6), (0.8, 0.1), (0.5, 0.5), (0.0, 0.99)]   # crop origin as a fraction of the slack


def add_affine_inputs(inp, B, H, W):
    angle = torch.tensor([[ANGLES[b % 4]] for b in range(B)])
    ratio = torch.tensor([[RATIOS[b % 4]] for b in range(B)])
    origin = torch.tensor([ORIGIN[b % 4] for b in range(B)])
    Rc, box = A.augmentation_geometry(inp[("K", 0)].cpu(), inp[("inv_K", 0)].cpu(), angle, ratio, origin, H, W)
    cpu = {k: v.cpu() for k, v in inp.items()}
    extra = {"Rc": Rc, "ratio_local": ratio, "angle": angle, "box": box}
    rec = (A.affine_transform(torch.ones(B, 1, H, W), angle, box) > 0).float()
    extra["valid_mask_rec"] = rec
    extra["valid_mask_cons"] = (A.depth_restore(rec, angle, box, torch.ones(B, 1)) > 0).float()
    for f in (-1, 0, 1):
        extra[("color_affine", f, 0)] = A.affine_transform(cpu[("color", f, 0)], angle, box)
    extra[("color_affine_aug", 0, 0)] = A.affine_transform(cpu[("color_aug", 0, 0)], angle, box)
    dev = inp[("K", 0)].device
    inp.update({k: v.to(dev) for k, v in extra.items()})
    return inp
