"""Golden vectors of the affine-augmentation transforms, produced by the UNMODIFIED reference's own methods on CPU
(build container only): Trainer.affine_transform (train.py:887-901) and the depth_restore tensors that
Trainer.compute_depth_consistency_loss_affine (train.py:903-921) hands to compute_SI_log_depth_loss, plus the gradient of
a fixed linear functional of them with respect to the affine depth.  Writes tests/golden/affine_golden.npz, consumed by
tests/test_affine.py."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import ref_harness  # noqa: E402
import net_fill  # noqa: E402

B, H, W = 3, 32, 64
T = ref_harness.import_reference(H, W, B)
import numpy as np  # noqa: E402
import torch  # noqa: E402

tr = ref_harness.make_trainer(T, B, H, W)
# angle / resize ratio / box the way datasets/mono_dataset.py:110-136 derives them (ratio in [1.2, 2.0], |angle| <= 10)
angle = torch.tensor([[7.25], [-3.5], [0.0]])
ratio = torch.tensor([[1.37], [1.9], [1.2]])
w0h0 = [(11, 5), (0, 13), (6, 0)]
box = []
for b in range(B):
    r = float(ratio[b, 0])
    box.append([round(w0h0[b][0] / r), round(w0h0[b][1] / r), round(W / r), round(H / r)])
box = torch.tensor(box)
inputs = {("angle"): angle, ("box"): box, ("ratio_local"): ratio,
          ("valid_mask_cons"): torch.ones(B, 1, H, W)}
img = net_fill.seeded_input((B, 3, H, W), 900)
img_affine = tr.affine_transform(img, inputs)

depth_affine = (net_fill.seeded_input((B, 1, H, W), 901) * 5 + 0.5).requires_grad_(True)
depth = net_fill.seeded_input((B, 1, H, W), 902) * 5 + 0.5
depth_fuse = net_fill.seeded_input((B, 1, H, W), 903) * 5 + 0.5
captured = []
orig = tr.compute_SI_log_depth_loss


def spy(pred, target, mask=None, beta=0.5):
    captured.append(pred)
    return orig(pred, target, mask, beta)


tr.compute_SI_log_depth_loss = spy
loss = tr.compute_depth_consistency_loss_affine(depth_affine, depth, depth_fuse, inputs)
restore = torch.cat(captured[0::2], 0)          # one per sample (each sample calls the SI-log loss twice)
weights = net_fill.seeded_input((B, 1, H, W), 904)
(restore * weights).sum().backward()
np.savez_compressed(os.path.join(HERE, "affine_golden.npz"), angle=angle.numpy(), box=box.numpy(), ratio=ratio.numpy(),
                    img=img.numpy(), img_affine=img_affine.detach().numpy(), depth_affine=depth_affine.detach().numpy(),
                    depth=depth.numpy(), depth_fuse=depth_fuse.numpy(), restore=restore.detach().numpy(),
                    weights=weights.numpy(), grad_depth_affine=depth_affine.grad.numpy(), loss_dc_affine=float(loss))
print("wrote affine_golden.npz: img_affine", tuple(img_affine.shape), "restore", tuple(restore.shape), "loss", float(loss))
