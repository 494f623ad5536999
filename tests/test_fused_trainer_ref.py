"""Build container only (needs /root/reference): the UNMODIFIED reference train.py, imported with this package's
`layers` / `networks` registered under the reference's module names, runs its own `process_batch` through
`FusedTrainer(Trainer)` and reproduces the losses the all-reference run produced (tests/golden/step_golden.json).

No GPU here, so the two CUDA-only pieces get test doubles: the fused loss group is evaluated by the C oracle (an
autograd.Function over oracle/f1_oracle.c) and `disp_to_depth` by its two-line formula.  Everything else -- the star
imports of train.py:18,20, model construction by the reference's names, predict_poses, the six loss groups, the SI-log
terms, the FusedTrainer overrides -- is the code that ships."""
import json
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree only exists in the build container")


class _OracleLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disp, P0, P1, tgt, s0, s1, inv_K, noise, mask, flags):
        from oracle import f1_oracle as O
        a = [None if t is None else t.detach().numpy() for t in (disp, tgt, s0, s1, inv_K, P0, P1, noise, mask)]
        out = O.f1_forward(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], flags=flags, full=False)
        ctx.a, ctx.idx, ctx.flags = a, out["idx"], flags
        return torch.tensor(float(out["loss"][0]), dtype=torch.float32)

    @staticmethod
    def backward(ctx, g):
        from oracle import f1_oracle as O
        a = ctx.a
        gd, gP0, gP1 = O.f1_backward(a[0], a[1], a[2], a[3], a[4], a[5], a[6], ctx.idx, a[8], float(g), flags=ctx.flags)
        return torch.from_numpy(gd), torch.from_numpy(gP0), torch.from_numpy(gP1), None, None, None, None, None, None, None


def _oracle_loss_fn(disp, tgt, s0, s1, inv_K, P0, P1, noise, mask, min_depth, max_depth, smooth_w, no_ssim, avg, noauto):
    flags = (1 if no_ssim else 0) | (2 if avg else 0) | (4 if noauto else 0)
    assert (min_depth, max_depth, smooth_w) == (0.1, 100.0, 1e-3)
    return _OracleLoss.apply(disp, P0, P1, tgt, s0, s1, inv_K, noise, mask, flags), None


def test_unmodified_train_py_runs_process_batch_through_fused_trainer(monkeypatch):
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import ref_harness
    import net_fill
    from mono_vifi_b200 import fused_trainer, layers, networks
    saved = {k: sys.modules.get(k) for k in ("layers", "networks", "train", "options", "utils", "kitti_utils", "datasets")}
    for k in saved:
        sys.modules.pop(k, None)
    try:
        ref_harness._install_shims()
        sys.modules["layers"], sys.modules["networks"] = layers, networks
        gold = json.load(open(os.path.join(HERE, "golden", "step_golden.json")))
        B, H, W = gold["B"], gold["H"], gold["W"]
        T = ref_harness.import_reference(H, W, B)          # `import train`: the reference's file, our layers / networks
        assert T.__file__.startswith(REF) and T.SSIM is layers.SSIM and T.monodepth2 is networks.monodepth2
        assert T.FusionModule is networks.FusionModule and T.IFRNet is networks.IFRNet
        FusedTrainer = fused_trainer.make_fused_trainer(T.Trainer, loss_fn=_oracle_loss_fn)
        assert fused_trainer.FusedTrainer.__mro__[1] is T.Trainer     # the lazy module attribute resolves too
        tr = FusedTrainer.__new__(FusedTrainer)
        import copy
        opt = copy.copy(T.opts)
        opt.height, opt.width, opt.batch_size, opt.use_affine = H, W, B, False
        opt.backbone, opt.fuse_model_type = "ResNet18", "shared_encoder"
        tr.opt, tr.device = opt, torch.device("cpu")
        torch.manual_seed(0)
        m = {}
        m["encoder"] = T.monodepth2.DepthEncoder(18, False)                       # train.py:142-146, by the reference's names
        m["depth"] = T.monodepth2.DepthDecoder(m["encoder"].num_ch_enc, range(1))
        m["encoder_mf"] = m["encoder"]
        m["depth_mf"] = copy.deepcopy(m["depth"])
        m["fusion_module"] = T.FusionModule(opt, m["encoder_mf"].num_ch_enc)
        m["pose_encoder"] = T.posenet.ResnetEncoder(18, False, num_input_images=2)
        m["pose"] = T.posenet.PoseDecoder(m["pose_encoder"].num_ch_enc, num_input_features=1, num_frames_to_predict_for=2)
        for name, mod in m.items():
            if name != "encoder_mf":
                net_fill.fill_(mod, scale=0.5 if name != "depth_mf" else 0.6)
                mod.train()
        tr.models = m
        tr.model_vfi_train = net_fill.fill_(T.IFRNet("small"), scale=0.7).eval()
        tr.ssim, tr.backproject_depth, tr.project_3d = T.SSIM(), T.BackprojectDepth(B, H, W), T.Project3D(B, H, W)

        def disp_to_depth(disp, min_depth, max_depth):   # layers.py:16-25 (the drop-in is a CUDA op)
            scaled = 1 / max_depth + (1 / min_depth - 1 / max_depth) * disp
            return scaled, 1 / scaled
        monkeypatch.setattr(T, "disp_to_depth", disp_to_depth)
        inp = {}
        for i, f in enumerate((-1, 0, 1)):
            inp[("color", f, 0)] = net_fill.seeded_input((B, 3, H, W), 500 + i)
            inp[("color_aug", f, 0)] = net_fill.seeded_input((B, 3, H, W), 510 + i)
        K = np.array([[0.58 * W, 0, 0.5 * W, 0], [0, 1.92 * H, 0.5 * H, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float32)
        inp[("K", 0)] = torch.from_numpy(np.repeat(K[None], B, 0).copy())
        inp[("inv_K", 0)] = torch.from_numpy(np.repeat(np.linalg.pinv(K)[None], B, 0).astype(np.float32).copy())
        torch.manual_seed(1)
        FusedTrainer.fused_groups = 0
        _, losses = tr.process_batch(inp)                  # train.py:698-885, unmodified
        assert FusedTrainer.fused_groups == 6              # every loss group went through the fused entry point
        g = gold["ResNet18"]
        assert abs(float(losses["loss_base"]) - g["loss_base"]) <= 1e-4 * g["loss_base"], (float(losses["loss_base"]), g["loss_base"])
        assert abs(float(losses["loss_dc"]) - g["loss_dc"]) <= 2e-4 * g["loss_dc"] + 1e-7, (float(losses["loss_dc"]), g["loss_dc"])
        losses["loss"].backward()
        gg = np.load(os.path.join(HERE, "golden", "step_grads_ResNet18.npz"))
        ref = {str(n): (s, a) for n, s, a in zip(gg["names"], gg["gsum"], gg["gabs"])}
        worst = 0.0
        for mname, mod in m.items():
            if mname == "encoder_mf":
                continue
            for pname, p in mod.named_parameters():
                s_ref, a_ref = ref["%s.%s" % (mname, pname)]
                if p.grad is None:
                    assert a_ref == 0.0
                    continue
                e = max(abs(float(p.grad.double().sum()) - s_ref), abs(float(p.grad.double().abs().sum()) - a_ref)) / max(a_ref, 1e-12)
                worst = max(worst, e)
        assert worst <= 2e-2, worst      # argmin ties broken by different noise draws move single pixels between sources
    finally:
        for k, v in saved.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v
        sys.path.remove(os.path.join(HERE, "golden"))
        if REF in sys.path:
            sys.path.remove(REF)
