"""Batched, sync-free affine-augmentation transforms (mono_vifi_b200/affine.py) against vectors produced by the
unmodified reference's own per-sample methods (tests/golden/gen_affine_golden.py: Trainer.affine_transform,
train.py:887-901, and the depth_restore tensors of Trainer.compute_depth_consistency_loss_affine, train.py:903-921)."""
import os

import numpy as np
import pytest
import torch

from mono_vifi_b200 import affine

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "affine_golden.npz"))
t = lambda k: torch.from_numpy(G[k])


def test_affine_transform_matches_reference():
    out = affine.affine_transform(t("img"), t("angle"), t("box"))
    # two bilinear resamplings of values in [0, 1): a few fp32 ulps of difference in how the lerps are associated
    np.testing.assert_allclose(out.numpy(), G["img_affine"], rtol=0, atol=2e-6)


def test_depth_restore_and_its_adjoint_match_reference():
    d = t("depth_affine").clone().requires_grad_(True)
    restore = affine.depth_restore(d, t("angle"), t("box"), t("ratio"))
    np.testing.assert_allclose(restore.detach().numpy(), G["restore"], rtol=2e-6, atol=2e-6)
    (restore * t("weights")).sum().backward()
    ref = G["grad_depth_affine"]
    np.testing.assert_allclose(d.grad.numpy(), ref, rtol=1e-5, atol=1e-6 * np.abs(ref).max())


def test_rotation_by_zero_and_full_box_are_identities():
    x = torch.rand(2, 3, 16, 24)
    zero = torch.zeros(2, 1)
    np.testing.assert_allclose(affine.rotate(x, zero).numpy(), x.numpy(), rtol=0, atol=1e-6)
    full = torch.tensor([[0, 0, 24, 16]] * 2)
    np.testing.assert_allclose(affine.crop_resize(x, full, (16, 24)).numpy(), x.numpy(), rtol=0, atol=1e-7)
    np.testing.assert_allclose(affine.shrink_paste(x, full).numpy(), x.numpy(), rtol=0, atol=1e-7)


@pytest.mark.parametrize("box", [(3, 2, 17, 11), (0, 0, 5, 4), (7, 5, 17, 11)])
def test_crop_resize_and_shrink_paste_match_torch_per_sample(box):
    x0, y0, w, h = box
    x = torch.rand(1, 2, 16, 24)
    ref = torch.nn.functional.interpolate(x[:, :, y0:y0 + h, x0:x0 + w], [16, 24], mode="bilinear", align_corners=False)
    np.testing.assert_allclose(affine.crop_resize(x, torch.tensor([box]), (16, 24)).numpy(), ref.numpy(), rtol=0, atol=1e-6)
    canvas = torch.zeros_like(x)
    canvas[:, :, y0:y0 + h, x0:x0 + w] = torch.nn.functional.interpolate(x, [h, w], mode="bilinear", align_corners=False)
    np.testing.assert_allclose(affine.shrink_paste(x, torch.tensor([box])).numpy(), canvas.numpy(), rtol=0, atol=1e-6)


def test_conjugate_pose_layout():
    torch.manual_seed(0)
    pose = torch.eye(4).repeat(2, 1, 1) + 0.01 * torch.randn(2, 4, 4)
    Rc = torch.eye(3).repeat(2, 1, 1) + 0.05 * torch.randn(2, 3, 3)
    out = affine.conjugate_pose(pose, Rc)
    ref = torch.zeros_like(pose)                                       # train.py:820-823
    ref[:, :3, :3] = torch.matmul(Rc, torch.matmul(pose[:, :3, :3], torch.inverse(Rc)))
    ref[:, :3, 3:4] = torch.matmul(Rc, pose[:, :3, 3:4])
    assert torch.equal(out, ref) and float(out[:, 3].abs().sum()) == 0.0


def _si_log_torch(pred, target, mask=None, beta=0.5):
    """test-only restatement of Trainer.compute_SI_log_depth_loss (train.py:924-941); the product path is the CUDA kernel"""
    if mask is None:
        mask = torch.ones_like(pred)
    mask = mask[:, 0]
    d = torch.log(pred[:, 0] + 1e-7) * mask - torch.log(target[:, 0] + 1e-7) * mask
    n = mask.sum(1).sum(1) + 1e-8
    return ((d ** 2).sum(1).sum(1) / n - beta * (d.sum(1).sum(1)) ** 2 / n ** 2).mean()


@pytest.mark.parametrize("use_affine", [False, True])
def test_depth_consistency_terms_of_the_step_match_reference_on_cpu(use_affine, monkeypatch):
    """Host-side wiring of process_batch's depth-consistency loss (train.py:806-812 and, with use_affine, 842, 863, 883):
    the networks, the frozen VFI model, the batched affine transforms and the pose / mask plumbing run on CPU against
    the loss_dc the unmodified reference computed (tests/golden/step_golden.json).  The photometric loss groups and the
    SI-log reduction are CUDA kernels (no CPU path): they are replaced by test doubles here and are compared with the
    reference in the GPU tests (so is disp_to_depth)."""
    import json
    import affine_inputs
    import net_fill
    from mono_vifi_b200 import networks as N, trainer as TR
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "step_golden.json")))
    B, H, W = gold["B"], gold["H"], gold["W"]
    opt = TR.Options(batch_size=B, height=H, width=W, backbone="ResNet18", multi_frame=True, fuse_model_type="shared_encoder",
                     use_affine=use_affine)
    torch.manual_seed(0)
    models = TR.build_models(opt, torch.device("cpu"))
    for name, mod in models.items():
        if name != "encoder_mf":
            net_fill.fill_(mod, scale=0.5 if name != "depth_mf" else 0.6)
        mod.train()
    vfi = net_fill.fill_(N.IFRNet("small"), scale=0.7).eval()
    inp = {}
    for i, f in enumerate((-1, 0, 1)):
        inp[("color", f, 0)] = net_fill.seeded_input((B, 3, H, W), 500 + i)
        inp[("color_aug", f, 0)] = net_fill.seeded_input((B, 3, H, W), 510 + i)
    K = np.array([[0.58 * W, 0, 0.5 * W, 0], [0, 1.92 * H, 0.5 * H, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float32)
    inp[("K", 0)] = torch.from_numpy(np.repeat(K[None], B, 0).copy())
    inp[("inv_K", 0)] = torch.from_numpy(np.repeat(np.linalg.pinv(K)[None], B, 0).astype(np.float32).copy())
    if use_affine:
        affine_inputs.add_affine_inputs(inp, B, H, W)
    seen = []

    def fake_group(opt_, disp, tgt, T0, T1, s0, s1, K_, iK_, mask_rec=None, noise=None):
        seen.append((tuple(disp.shape), tuple(T0.shape), mask_rec is not None))
        return disp.mean() * 0 + T0.sum() * 0 + T1.sum() * 0, None

    monkeypatch.setattr(TR, "loss_group", fake_group)
    monkeypatch.setattr(TR.L, "si_log_depth_loss", _si_log_torch)

    def disp_to_depth(disp, min_depth, max_depth):   # layers.py:16-25, test double of the CUDA op
        scaled = 1 / max_depth + (1 / min_depth - 1 / max_depth) * disp
        return scaled, 1 / scaled

    monkeypatch.setattr(TR.L, "disp_to_depth", disp_to_depth)
    out = TR.multi_frame_losses(models, vfi, inp, opt)
    g = gold["ResNet18_affine" if use_affine else "ResNet18"]
    assert abs(float(out["loss_dc"]) - g["loss_dc"]) <= 2e-4 * g["loss_dc"] + 1e-7, (float(out["loss_dc"]), g["loss_dc"])
    assert len(seen) == (9 if use_affine else 6) and [s[2] for s in seen] == [False] * 6 + [True] * (3 if use_affine else 0)
    out["loss"].backward()   # the affine terms reach the single-frame networks and, through the conjugated poses, the pose nets
    assert all(torch.isfinite(p.grad).all() for p in models["depth"].parameters() if p.grad is not None)
