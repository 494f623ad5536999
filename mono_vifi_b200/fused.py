"""Fused view-synthesis + photometric loss (kernel F1) as a torch.autograd.Function.

Replaces, per loss group, Trainer.generate_images_pred x2 + compute_reprojection_loss x4 +
compute_losses_base of the reference (train.py:956-1051); see include/monovifi_b200.h.
"""
import torch

from . import _lib

_workspaces = {}

# launch accounting for bench.py: number of F1 kernel launches, and -- when `timing` is a list -- a pair of CUDA
# events around every launch (recorded on the launching stream), tagged "f1_fwd" / "f1_bwd"
launches = {"f1_fwd": 0, "f1_bwd": 0}
timing = None


def _timed(tag, fn):
    launches[tag] += 1
    if timing is None:
        return fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = fn()
    e1.record()
    timing.append((tag, e0, e1))
    return rc


def _ptr(t):
    return None if t is None else t.data_ptr()


def _prep(t, shape=None):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("mono_vifi_b200 ops run on CUDA tensors only (no CPU fallback)")
    if t.dtype != torch.float32:
        t = t.float()
    t = t.contiguous()
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError("expected shape %s, got %s" % (tuple(shape), tuple(t.shape)))
    return t


def workspace(device, B):
    """One zero-initialised, self-cleaning workspace per (device, stream); see mvf_workspace_init."""
    stream = torch.cuda.current_stream(device)
    key = (device.index, stream.cuda_stream)
    need = _lib.lib().mvf_f1_workspace_bytes(B)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(max(need, 4096), dtype=torch.uint8, device=device)
        _lib.check(_lib.lib().mvf_workspace_init(ws.data_ptr(), ws.numel(), stream.cuda_stream), "workspace_init")
        _workspaces[key] = ws
    return ws, stream.cuda_stream


def flags_of(no_ssim=False, avg_reprojection=False, disable_automasking=False):
    return (_lib.NO_SSIM if no_ssim else 0) | (_lib.AVG_REPROJECTION if avg_reprojection else 0) | \
        (_lib.DISABLE_AUTOMASKING if disable_automasking else 0)


def f1_forward_raw(disp, tgt, src0, src1, inv_K, P0, P1, noise=None, mask_rec=None, min_depth=0.1, max_depth=100.0,
                   smooth_w=1e-3, flags=0, debug=False):
    """Launches mvf_f1_forward.  Returns dict(loss[4], stats, idx, [x0y0, warp0, warp1, to_optimise])."""
    B, _, H, W = disp.shape
    dev = disp.device
    disp = _prep(disp, (B, 1, H, W))
    tgt, src0, src1 = _prep(tgt, (B, 3, H, W)), _prep(src0, (B, 3, H, W)), _prep(src1, (B, 3, H, W))
    inv_K, P0, P1 = _prep(inv_K, (B, 4, 4)), _prep(P0, (B, 3, 4)), _prep(P1, (B, 3, 4))
    am, avg = not (flags & _lib.DISABLE_AUTOMASKING), bool(flags & _lib.AVG_REPROJECTION)
    nid = (1 if avg else 2) if am else 0
    noise = _prep(noise, (B, nid, H, W)) if (noise is not None and nid) else None
    mask_rec = _prep(mask_rec, (B, 1, H, W))
    out = {"loss": torch.empty(4, device=dev), "stats": torch.empty(B, 4, device=dev),
           "idx": torch.empty(B, H, W, dtype=torch.uint8, device=dev)}
    x0y0 = w0 = w1 = topt = None
    if debug:
        x0y0 = torch.empty(2, 2, B, H, W, dtype=torch.int32, device=dev)
        w0, w1 = torch.empty(B, 3, H, W, device=dev), torch.empty(B, 3, H, W, device=dev)
        topt = torch.empty(B, H, W, device=dev)
        out.update(x0y0=x0y0, warp0=w0, warp1=w1, to_optimise=topt)
    prm = _lib.f1_params(B, H, W, min_depth, max_depth, smooth_w, flags)
    ws, stream = workspace(dev, B)
    rc = _timed("f1_fwd", lambda: _lib.lib().mvf_f1_forward(
        prm, _ptr(disp), _ptr(tgt), _ptr(src0), _ptr(src1), _ptr(inv_K), _ptr(P0), _ptr(P1), _ptr(noise), _ptr(mask_rec),
        _ptr(out["loss"]), _ptr(out["stats"]), _ptr(out["idx"]), _ptr(x0y0), _ptr(w0), _ptr(w1), _ptr(topt),
        ws.data_ptr(), ws.numel(), stream))
    _lib.check(rc, "mvf_f1_forward")
    out["_saved"] = (disp, tgt, src0, src1, inv_K, P0, P1, mask_rec)
    return out


def f1_backward_raw(saved, idx, stats, gout=None, min_depth=0.1, max_depth=100.0, smooth_w=1e-3, flags=0):
    disp, tgt, src0, src1, inv_K, P0, P1, mask_rec = saved
    B, _, H, W = disp.shape
    dev = disp.device
    g_disp = torch.empty_like(disp)
    g_P0, g_P1 = torch.empty(B, 3, 4, device=dev), torch.empty(B, 3, 4, device=dev)
    if gout is not None:
        gout = _prep(gout.reshape(1))
    prm = _lib.f1_params(B, H, W, min_depth, max_depth, smooth_w, flags)
    ws, stream = workspace(dev, B)
    rc = _timed("f1_bwd", lambda: _lib.lib().mvf_f1_backward(
        prm, _ptr(disp), _ptr(tgt), _ptr(src0), _ptr(src1), _ptr(inv_K), _ptr(P0), _ptr(P1), _ptr(mask_rec), _ptr(idx),
        _ptr(stats), _ptr(gout), _ptr(g_disp), _ptr(g_P0), _ptr(g_P1), ws.data_ptr(), ws.numel(), stream))
    _lib.check(rc, "mvf_f1_backward")
    return g_disp, g_P0, g_P1


class _FusedPhotometricLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disp, P0, P1, tgt, src0, src1, inv_K, noise, mask_rec, min_depth, max_depth, smooth_w, flags):
        out = f1_forward_raw(disp, tgt, src0, src1, inv_K, P0, P1, noise, mask_rec, min_depth, max_depth, smooth_w,
                             flags)
        ctx.saved = out["_saved"]
        ctx.idx, ctx.stats = out["idx"], out["stats"]
        ctx.cfg = (min_depth, max_depth, smooth_w, flags)
        ctx.mark_non_differentiable(out["idx"])
        return out["loss"][0], out["idx"], out["loss"]

    @staticmethod
    def backward(ctx, g_loss, _g_idx, _g_parts):
        g_disp, g_P0, g_P1 = f1_backward_raw(ctx.saved, ctx.idx, ctx.stats, g_loss, *ctx.cfg)
        return (g_disp, g_P0, g_P1) + (None,) * 10


def fused_photometric_loss(disp, img_tgt, img_src0, img_src1, inv_K, P0, P1, noise=None, mask_rec=None,
                           min_depth=0.1, max_depth=100.0, smooth_w=1e-3, no_ssim=False, avg_reprojection=False,
                           disable_automasking=False):
    """loss, auto_mask = one loss group of the reference (train.py:956-1051).

    disp [B,1,H,W] (requires grad), P0/P1 = (K @ T)[:, :3] [B,3,4] (require grad; autograd finishes K@T and
    the Rodrigues map), images [B,3,H,W], inv_K [B,4,4].  `noise` is the tensor train.py:1023 draws with
    torch.randn ([B,2,H,W]; [B,1,H,W] with avg_reprojection); None = no tie-break noise.
    Returns (loss 0-dim, auto_mask [B,1,H,W] float: argmin picked a warped candidate, train.py:1036).
    """
    flags = flags_of(no_ssim, avg_reprojection, disable_automasking)
    loss, idx, _ = _FusedPhotometricLoss.apply(disp, P0, P1, img_tgt, img_src0, img_src1, inv_K, noise, mask_rec,
                                               float(min_depth), float(max_depth), float(smooth_w), flags)
    nid = 0 if disable_automasking else (1 if avg_reprojection else 2)
    auto_mask = (idx >= nid).unsqueeze(1).float() if nid else None  # train.py:1035-1038
    return loss, auto_mask
