"""Drop-in for the reference's `layers.py` (same names, signatures, outputs and autograd behaviour), with the
arithmetic in libmonovifi_b200.so.  `from layers import *` in the reference's train.py / networks resolves to
this module when `mono_vifi_b200/dropin` is ahead of the reference on sys.path.

CUDA tensors only: there is no CPU fallback (a CPU tensor raises).  Reference lines are cited per symbol.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .fused import _prep, _ptr, workspace
from . import conv as conv_mod
from . import decoder_ops
from .conv import Conv2d


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _disp_consts(min_depth, max_depth):
    min_disp = 1 / max_depth
    max_disp = 1 / min_depth
    return float(min_disp), float(max_disp - min_disp)


class _DispToDepth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disp, min_disp, rng):
        d = _prep(disp)
        sd, depth = torch.empty_like(d), torch.empty_like(d)
        _lib.check(_lib.lib().mvf_disp_to_depth_fwd(_ptr(d), _ptr(sd), _ptr(depth), d.numel(), min_disp, rng, _stream(d)),
                   "disp_to_depth_fwd")
        ctx.save_for_backward(d)
        ctx.cfg = (min_disp, rng)
        return sd, depth

    @staticmethod
    def backward(ctx, g_sd, g_depth):
        (d,) = ctx.saved_tensors
        g = torch.empty_like(d)
        g_sd = None if g_sd is None else _prep(g_sd)
        g_depth = None if g_depth is None else _prep(g_depth)
        _lib.check(_lib.lib().mvf_disp_to_depth_bwd(_ptr(d), _ptr(g_sd), _ptr(g_depth), _ptr(g), d.numel(), ctx.cfg[0],
                                                    ctx.cfg[1], _stream(d)), "disp_to_depth_bwd")
        return g, None, None


def disp_to_depth(disp, min_depth, max_depth):
    """layers.py:16-25 -> (scaled_disp, depth)"""
    min_disp, rng = _disp_consts(min_depth, max_depth)
    return _DispToDepth.apply(disp, min_disp, rng)


def transformation_from_parameters(axisangle, translation, invert=False):
    """layers.py:28-45.  [B,1,3] x2 -> [B,4,4].  On CUDA: one kernel forward, one backward (mvf_pose_matrix_*) instead of
    the ~40 small launches of the op-by-op form below, which CPU tensors still take."""
    if axisangle.is_cuda:
        from . import warp_ops
        return warp_ops.pose_matrix(axisangle, translation, invert)
    R = rot_from_axisangle(axisangle)
    t = translation.clone()
    if invert:
        R = R.transpose(1, 2)
        t = t * -1
    T = get_translation_matrix(t)
    return torch.matmul(R, T) if invert else torch.matmul(T, R)


def get_translation_matrix(translation_vector):
    """layers.py:48-63"""
    B = translation_vector.shape[0]
    t = translation_vector.contiguous().view(B, 3, 1)
    eye = torch.eye(4, device=t.device, dtype=t.dtype).unsqueeze(0).repeat(B, 1, 1)
    return torch.cat([eye[:, :, :3], torch.cat([t, eye[:, 3:, 3:]], 1)], 2)


def rot_from_axisangle(vec):
    """layers.py:66-103 (Rodrigues); vec [B,1,3] -> [B,4,4].  Same op sequence, assembled without in-place writes."""
    angle = torch.norm(vec, 2, 2, True)
    axis = vec / (angle + 1e-7)
    ca, sa = torch.cos(angle), torch.sin(angle)
    C = 1 - ca
    x, y, z = axis[..., 0].unsqueeze(1), axis[..., 1].unsqueeze(1), axis[..., 2].unsqueeze(1)
    xs, ys, zs = x * sa, y * sa, z * sa
    xC, yC, zC = x * C, y * C, z * C
    xyC, yzC, zxC = x * yC, y * zC, z * xC
    zero, one = torch.zeros_like(ca), torch.ones_like(ca)
    rows = [torch.cat([x * xC + ca, xyC - zs, zxC + ys, zero], 2),
            torch.cat([xyC + zs, y * yC + ca, yzC - xs, zero], 2),
            torch.cat([zxC - ys, yzC + xs, z * zC + ca, zero], 2),
            torch.cat([zero, zero, zero, one], 2)]
    return torch.cat(rows, 1)


class ConvBlock(nn.Module):
    """layers.py:106-119: Conv3x3 + ELU"""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = Conv3x3(in_channels, out_channels)
        self.nonlin = nn.ELU()

    def forward(self, x):
        if self.conv.fast_path(x):
            return self.conv.forward_padded(decoder_ops.upcat_pad(x), act="elu")
        return self.nonlin(self.conv(x))

    def forward_upcat(self, x, skip=None, upsample=True):
        """ConvBlock(cat([upsample(x), skip], 1)) with the upsample + concat + reflection pad done in one kernel
        (what monodepth2.py:86-93 feeds to upconv(i,1))."""
        if self.conv.fast_path(x) and decoder_ops.usable(x, skip, upsample):
            return self.conv.forward_padded(decoder_ops.upcat_pad(x, skip, upsample), act="elu")
        x = upsample_fn(x) if upsample else x
        if skip is not None:
            x = torch.cat([x, skip], 1)
        return self.forward(x)


class Conv3x3(nn.Module):
    """layers.py:122-139: reflection (or zero) pad 1 + 3x3 conv; state_dict key `conv.weight` / `conv.bias`."""

    def __init__(self, in_channels, out_channels, use_refl=True):
        super().__init__()
        self.pad = nn.ReflectionPad2d(1) if use_refl else nn.ZeroPad2d(1)
        self.conv = Conv2d(int(in_channels), int(out_channels), 3)

    def fast_path(self, x):
        """channels-last fused pad (+ tcgen05 convolution with the activation in its epilogue)"""
        return (x.is_cuda and isinstance(self.pad, nn.ReflectionPad2d) and conv_mod.get_backend() == "tcgen05"
                and decoder_ops.usable(x))

    def forward_padded(self, xpad, act=None):
        return conv_mod.conv2d(xpad, self.conv.weight, self.conv.bias, 1, 0, act=act)

    def forward(self, x):
        if self.fast_path(x):
            return self.forward_padded(decoder_ops.upcat_pad(x))
        return self.conv(self.pad(x))


class Conv1x1(nn.Module):
    """layers.py:142-151"""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = Conv2d(int(in_channels), int(out_channels), kernel_size=1, stride=1)

    def forward(self, x):
        return self.conv(x)


class ConvBlock1x1(nn.Module):
    """layers.py:154-167"""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = Conv1x1(in_channels, out_channels)
        self.nonlin = nn.ELU()

    def forward(self, x):
        return self.nonlin(self.conv(x))


class _Backproject(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, inv_K, B, H, W):
        d = _prep(depth.reshape(B, 1, H, W))
        k = _prep(inv_K, (B, 4, 4))
        out = torch.empty(B, 4, H * W, device=d.device)
        _lib.check(_lib.lib().mvf_backproject_fwd(_ptr(d), _ptr(k), _ptr(out), B, H, W, _stream(d)), "backproject_fwd")
        ctx.save_for_backward(k)
        ctx.dims = (B, H, W, depth.shape)
        return out

    @staticmethod
    def backward(ctx, g):
        (k,) = ctx.saved_tensors
        B, H, W, shape = ctx.dims
        g = _prep(g)
        gd = torch.empty(B, 1, H, W, device=g.device)
        _lib.check(_lib.lib().mvf_backproject_bwd(_ptr(g), _ptr(k), _ptr(gd), B, H, W, _stream(g)), "backproject_bwd")
        return gd.reshape(shape), None, None, None, None


class BackprojectDepth(nn.Module):
    """layers.py:168-197.  Keeps the reference's non-trainable Parameters (id_coords, ones, pix_coords) so that
    state_dicts and `.to(device)` behave identically; the kernel recomputes the pixel grid instead of reading it."""

    def __init__(self, batch_size, height, width):
        super().__init__()
        self.batch_size, self.height, self.width = batch_size, height, width
        meshgrid = np.meshgrid(range(self.width), range(self.height), indexing='xy')
        id_coords = np.stack(meshgrid, axis=0).astype(np.float32)
        self.id_coords = nn.Parameter(torch.from_numpy(id_coords), requires_grad=False)
        self.ones = nn.Parameter(torch.ones(self.batch_size, 1, self.height * self.width), requires_grad=False)
        pix = torch.unsqueeze(torch.stack([self.id_coords[0].view(-1), self.id_coords[1].view(-1)], 0), 0)
        pix = pix.repeat(batch_size, 1, 1)
        self.pix_coords = nn.Parameter(torch.cat([pix, self.ones], 1), requires_grad=False)

    def forward(self, depth, inv_K):
        return _Backproject.apply(depth, inv_K, self.batch_size, self.height, self.width)


def matmul_KT(K, T):
    """`torch.matmul(K, T)` of layers.py:212 for [B,4,4] operands, written as separately rounded multiply / add
    steps in k order.  That is what torch's CPU bmm produces for this shape (no FMA, verified bit for bit against
    the golden vectors in tests/test_networks_api.py), and -- unlike cuBLAS -- it gives the same bits on the GPU, so
    the sampling grid downstream is bit-identical to the CPU reference's.  Differentiable (plain torch ops)."""
    P = K[:, :, 0:1] * T[:, 0:1, :]
    for j in range(1, 4):
        P = P + K[:, :, j:j + 1] * T[:, j:j + 1, :]
    return P


class _Project(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, P, B, H, W, eps):
        pts = _prep(points, (B, 4, H * W))
        P = _prep(P, (B, 3, 4))
        grid = torch.empty(B, H, W, 2, device=pts.device)
        _lib.check(_lib.lib().mvf_project_fwd(_ptr(pts), _ptr(P), _ptr(grid), B, H, W, eps, _stream(pts)), "project_fwd")
        ctx.save_for_backward(pts, P)
        ctx.dims = (B, H, W, eps)
        return grid

    @staticmethod
    def backward(ctx, g):
        pts, P = ctx.saved_tensors
        B, H, W, eps = ctx.dims
        g = _prep(g)
        g_pts, g_P = torch.empty_like(pts), torch.empty_like(P)
        ws, stream = workspace(g.device, B)
        _lib.check(_lib.lib().mvf_project_bwd(_ptr(pts), _ptr(P), _ptr(g), _ptr(g_pts), _ptr(g_P), ws.data_ptr(),
                                              ws.numel(), B, H, W, eps, stream), "project_bwd")
        return g_pts, g_P, None, None, None, None


class Project3D(nn.Module):
    """layers.py:200-222.  `K @ T` (layers.py:212) goes through matmul_KT so its bits match the reference's
    (exactness recipe); the [3,4] x [4,HW] product, divide and normalise run in one kernel."""

    def __init__(self, batch_size, height, width, eps=1e-7):
        super().__init__()
        self.batch_size, self.height, self.width, self.eps = batch_size, height, width, eps

    def forward(self, points, K, T):
        P = matmul_KT(K, T)[:, :3, :]
        return _Project.apply(points, P, self.batch_size, self.height, self.width, float(self.eps))


def upsample_fn(x):
    return F.interpolate(x, scale_factor=2, mode="nearest")


def upsample(x, scale_factor=2, mode="nearest"):
    """layers.py:225-228 (mode="bilinear": the Lite-Mono decoder, LiteMono.py:495,502 -> mvf_resize_bilinear_* on CUDA)"""
    if mode == "bilinear" and x.is_cuda:
        from . import warp_ops
        return warp_ops.resize_bilinear(x, scale_factor=scale_factor, align_corners=False)
    return F.interpolate(x, scale_factor=scale_factor, mode=mode)


class _SmoothLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disp, img):
        B, _, H, W = disp.shape
        d, im = _prep(disp, (B, 1, H, W)), _prep(img, (B, 3, H, W))
        out = torch.empty(1, device=d.device)
        ws, stream = workspace(d.device, B)
        _lib.check(_lib.lib().mvf_smooth_loss_fwd(_ptr(d), _ptr(im), _ptr(out), ws.data_ptr(), ws.numel(), B, H, W,
                                                  stream), "smooth_loss_fwd")
        ctx.save_for_backward(d, im)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        d, im = ctx.saved_tensors
        B, _, H, W = d.shape
        gd = torch.empty_like(d)
        g = _prep(g.reshape(1))
        _lib.check(_lib.lib().mvf_smooth_loss_bwd(_ptr(d), _ptr(im), _ptr(g), _ptr(gd), B, H, W, _stream(d)),
                   "smooth_loss_bwd")
        return gd, None


def get_smooth_loss(disp, img):
    """layers.py:231-242.  Gradient flows to `disp` only (the image is data everywhere the reference calls it)."""
    if img.requires_grad:
        raise NotImplementedError("get_smooth_loss: gradient w.r.t. the image is not implemented (never used by train.py)")
    if img.shape[1] != 3 or disp.shape[1] != 1:
        raise ValueError("get_smooth_loss expects disp [B,1,H,W] and img [B,3,H,W]")
    return _SmoothLoss.apply(disp, img)


class _SSIM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        B, C, H, W = x.shape
        x, y = _prep(x), _prep(y, (B, C, H, W))
        out = torch.empty_like(x)
        _lib.check(_lib.lib().mvf_ssim_fwd(_ptr(x), _ptr(y), _ptr(out), B * C, H, W, _stream(x)), "ssim_fwd")
        ctx.save_for_backward(x, y)
        return out

    @staticmethod
    def backward(ctx, g):
        x, y = ctx.saved_tensors
        B, C, H, W = x.shape
        g = _prep(g)
        scratch = torch.empty(3, B * C, H, W, device=x.device)
        gx = gy = None
        if ctx.needs_input_grad[0]:
            gx = torch.empty_like(x)
            _lib.check(_lib.lib().mvf_ssim_bwd(_ptr(x), _ptr(y), _ptr(g), _ptr(gx), _ptr(scratch), B * C, H, W,
                                               _stream(x)), "ssim_bwd")
        if ctx.needs_input_grad[1]:
            gy = torch.empty_like(y)
            _lib.check(_lib.lib().mvf_ssim_bwd(_ptr(y), _ptr(x), _ptr(g), _ptr(gy), _ptr(scratch), B * C, H, W,
                                               _stream(x)), "ssim_bwd")
        return gx, gy


class SSIM(nn.Module):
    """layers.py:261-290: (1 - SSIM)/2 clamped to [0,1], 3x3 mean windows over reflection-padded inputs."""

    def __init__(self):
        super().__init__()
        self.C1 = 0.01 ** 2
        self.C2 = 0.03 ** 2

    def forward(self, x, y):
        return _SSIM.apply(x, y)


class _SILog(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, mask, beta):
        B = pred.shape[0]
        p, t = _prep(pred), _prep(target)
        m = None if mask is None else _prep(mask.expand_as(pred) if mask.shape != pred.shape else mask)
        HW = p.numel() // B
        loss, stats = torch.empty(1, device=p.device), torch.empty(B, 2, device=p.device)
        ws, stream = workspace(p.device, B)
        _lib.check(_lib.lib().mvf_si_log_fwd(_ptr(p), _ptr(t), _ptr(m), _ptr(loss), _ptr(stats), ws.data_ptr(),
                                             ws.numel(), B, HW, beta, stream), "si_log_fwd")
        ctx.save_for_backward(p, t, stats)
        ctx.mask, ctx.beta = m, beta
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        p, t, stats = ctx.saved_tensors
        B = p.shape[0]
        HW = p.numel() // B
        gp = torch.empty_like(p) if ctx.needs_input_grad[0] else None
        gt = torch.empty_like(t) if ctx.needs_input_grad[1] else None
        g = _prep(g.reshape(1))
        _lib.check(_lib.lib().mvf_si_log_bwd(_ptr(p), _ptr(t), _ptr(ctx.mask), _ptr(stats), _ptr(g), _ptr(gp), _ptr(gt),
                                             B, HW, ctx.beta, _stream(p)), "si_log_bwd")
        return gp, gt, None, None


def si_log_depth_loss(pred, target, mask=None, beta=0.5):
    """Trainer.compute_SI_log_depth_loss, train.py:924-941 (scale-invariant log depth consistency)."""
    return _SILog.apply(pred, target, mask, float(beta))


def compute_depth_errors(gt, pred):
    """layers.py:293-311 (evaluation metrics; plain torch, not on the training path)."""
    thresh = torch.max((gt / pred), (pred / gt))
    a1 = (thresh < 1.25).float().mean()
    a2 = (thresh < 1.25 ** 2).float().mean()
    a3 = (thresh < 1.25 ** 3).float().mean()
    rmse = torch.sqrt(((gt - pred) ** 2).mean())
    rmse_log = torch.sqrt(((torch.log(gt) - torch.log(pred)) ** 2).mean())
    abs_rel = torch.mean(torch.abs(gt - pred) / gt)
    sq_rel = torch.mean((gt - pred) ** 2 / gt)
    return abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3
