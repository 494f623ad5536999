"""`FusedTrainer(Trainer)`: the reference's own Trainer (train.py, unmodified) with its three view-synthesis / loss
methods routed to the fused kernel -- SURVEY.md 8(b).

The reference computes one loss group as

    a = self.generate_images_pred(disp, pose_a, img_src_a, K, inv_K)          train.py:956-971
    b = self.generate_images_pred(disp, pose_b, img_src_b, K, inv_K)
    loss, auto_mask = self.compute_losses_base(disp, img_tgt, [a, b], [img_src_a, img_src_b], mask_rec)   train.py:987-1051

`generate_images_pred` here does not warp anything: it returns a `PendingWarp` that remembers its arguments, and
`compute_losses_base`, which is where the warped images are consumed, launches ONE fused kernel for the whole group
(back-projection, projection, border-mode bilinear sampling of both sources, 4 x (SSIM + L1), tie-break noise, per-pixel
min, mask, mean, edge-aware smoothness).  `process_batch`, `predict_poses`, the affine branch, the optimiser loop -- all
of train.py -- run unchanged.  A caller that needs the warped tensor itself (anything that treats the return value as a
tensor) gets it: PendingWarp materialises through the stand-alone drop-in modules on first tensor use.

    import sys
    from mono_vifi_b200 import layers, networks
    sys.modules["layers"], sys.modules["networks"] = layers, networks       # train.py:18,20 star-imports
    import train                                                            # the unmodified reference
    from mono_vifi_b200.fused_trainer import make_fused_trainer
    FusedTrainer = make_fused_trainer(train.Trainer)
    FusedTrainer(train.opts).train()
"""
import torch

from . import layers as L


class PendingWarp:
    """What generate_images_pred returns: the arguments of one source's warp (train.py:956-971), not yet evaluated."""

    def __init__(self, trainer, disp_tgt, pose, img_src, K, inv_K):
        self.trainer, self.disp_tgt, self.pose, self.img_src, self.K, self.inv_K = trainer, disp_tgt, pose, img_src, K, inv_K
        self._tensor = None

    def tensor(self):
        """the warped image itself, through the stand-alone modules (for callers other than compute_losses_base)"""
        if self._tensor is None:
            t = self.trainer
            disp = self.disp_tgt[("disp", 0)]
            _, depth = L.disp_to_depth(disp, t.opt.min_depth, t.opt.max_depth)
            pix = t.project_3d(t.backproject_depth(depth, self.inv_K), self.K, self.pose)
            self._tensor = torch.nn.functional.grid_sample(self.img_src, pix, padding_mode="border", align_corners=True)
        return self._tensor

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        conv = lambda a: a.tensor() if isinstance(a, PendingWarp) else a
        return func(*[conv(a) for a in args], **{k: conv(v) for k, v in (kwargs or {}).items()})

    def __getattr__(self, name):      # .shape, .detach(), ... : behave like the tensor
        return getattr(self.tensor(), name)


def fused_loss_group(trainer, disp_tgt, img_tgt, pending, mask_rec=None, loss_fn=None):
    """compute_losses_base of two PendingWarps of the same disparity: one fused launch.  `loss_fn` defaults to
    fused.fused_photometric_loss (the CUDA kernel); tests pass the oracle-backed double on hosts without a GPU."""
    if loss_fn is None:
        from .fused import fused_photometric_loss as loss_fn
    opt = trainer.opt
    disp = disp_tgt[("disp", 0)]
    a, b = pending
    P0 = L.matmul_KT(a.K, a.pose)[:, :3, :]          # layers.py:212 stays a torch matmul: the reference's bits
    P1 = L.matmul_KT(b.K, b.pose)[:, :3, :]
    noise = None
    if not opt.disable_automasking:
        B, _, H, W = disp.shape
        noise = torch.randn(B, 1 if opt.avg_reprojection else 2, H, W, device=disp.device)   # train.py:1023
    return loss_fn(disp, img_tgt, a.img_src, b.img_src, a.inv_K, P0, P1, noise, mask_rec, opt.min_depth, opt.max_depth,
                   opt.disparity_smoothness, opt.no_ssim, opt.avg_reprojection, opt.disable_automasking)


def make_fused_trainer(Trainer, loss_fn=None):
    """Trainer: the reference's train.Trainer class (imported by the caller).  Returns the subclass."""

    class FusedTrainer(Trainer):
        fused_groups = 0     # loss groups that went through the fused kernel (diagnostics / tests)

        def generate_images_pred(self, disp_tgt, pose_tgt_src, img_src, K, inv_K):
            return PendingWarp(self, disp_tgt, pose_tgt_src, img_src, K, inv_K)

        def compute_losses_base(self, disp_tgt, img_tgt, imgs_src_tgt, imgs_src, mask_rec=None):
            fusable = (len(imgs_src_tgt) == 2 and all(isinstance(w, PendingWarp) for w in imgs_src_tgt) and
                       all(w.disp_tgt is disp_tgt for w in imgs_src_tgt) and
                       all(w.img_src is s for w, s in zip(imgs_src_tgt, imgs_src)))
            if not fusable:   # any other calling pattern: the reference's own arithmetic on the materialised warps
                warped = [w.tensor() if isinstance(w, PendingWarp) else w for w in imgs_src_tgt]
                return super().compute_losses_base(disp_tgt, img_tgt, warped, imgs_src, mask_rec)
            type(self).fused_groups += 1
            return fused_loss_group(self, disp_tgt, img_tgt, imgs_src_tgt, mask_rec, loss_fn)

        def compute_reprojection_loss(self, pred, target):
            if isinstance(pred, PendingWarp):
                pred = pred.tensor()
            return super().compute_reprojection_loss(pred, target)

    FusedTrainer.__qualname__ = FusedTrainer.__name__ = "FusedTrainer"
    return FusedTrainer


def __getattr__(name):
    """`from mono_vifi_b200.fused_trainer import FusedTrainer` works once the reference's `train` module is imported."""
    if name == "FusedTrainer":
        import sys
        train = sys.modules.get("train")
        if train is None or not hasattr(train, "Trainer"):
            raise ImportError("import the reference's train.py (with the drop-in layers / networks in sys.modules) first")
        return make_fused_trainer(train.Trainer)
    raise AttributeError(name)
