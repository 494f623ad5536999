"""ctypes/numpy front-end of the CPU oracle (oracle/f1_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product path (mono_vifi_b200) never does.  Each function cites the reference lines it
restates in f1_oracle.c.  Parity pin: tests/test_oracle_golden.py (golden vectors generated from the
unmodified reference by tests/golden/gen_golden.py).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmvf_oracle.so")

NO_SSIM, AVG_REPROJECTION, DISABLE_AUTOMASKING = 1, 2, 4


class Params(ctypes.Structure):
    _fields_ = [("B", ctypes.c_int), ("H", ctypes.c_int), ("W", ctypes.c_int),
                ("min_disp", ctypes.c_float), ("disp_range", ctypes.c_float),
                ("smooth_w", ctypes.c_float), ("flags", ctypes.c_int)]


def build(force=False):
    src = os.path.join(_HERE, "f1_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libmvf_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.mvfo_smooth_loss.restype = ctypes.c_double
        _lib.mvfo_si_log_loss.restype = ctypes.c_double
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def make_params(B, H, W, min_depth=0.1, max_depth=100.0, smooth_w=1e-3, flags=0):
    # python-double arithmetic then one cast, exactly what layers.py:21-23 hands to torch
    min_disp = 1 / max_depth
    max_disp = 1 / min_depth
    return Params(B, H, W, np.float32(min_disp), np.float32(max_disp - min_disp), np.float32(smooth_w), flags)


def disp_to_depth(disp, min_depth=0.1, max_depth=100.0):
    """layers.py:16-25"""
    disp, dp = _f(disp)
    sd = np.empty_like(disp)
    depth = np.empty_like(disp)
    p = make_params(1, 1, 1, min_depth, max_depth)
    lib().mvfo_disp_to_depth(dp, _p(sd), _p(depth), ctypes.c_long(disp.size),
                             ctypes.c_float(p.min_disp), ctypes.c_float(p.disp_range))
    return sd, depth


def backproject(depth, inv_K):
    """layers.py:192-197 -> [B,4,H*W]"""
    depth, dp = _f(depth)
    inv_K, kp = _f(inv_K)
    B, _, H, W = depth.shape
    out = np.empty((B, 4, H * W), np.float32)
    lib().mvfo_backproject(dp, kp, _p(out), B, H, W)
    return out


def project(points, P, H, W):
    """layers.py:211-222 with P = (K@T)[:, :3] -> grid [B,H,W,2]"""
    points, pp = _f(points)
    P, Pp = _f(P)
    B = points.shape[0]
    out = np.empty((B, H, W, 2), np.float32)
    lib().mvfo_project(pp, Pp, _p(out), B, H, W)
    return out


def grid_sample(img, grid):
    """train.py:966-969 (bilinear, border, align_corners=True) -> (out, x0, y0)"""
    img, ip = _f(img)
    grid, gp = _f(grid)
    B, C, H, W = img.shape
    out = np.empty_like(img)
    x0 = np.empty((B, H, W), np.int32)
    y0 = np.empty((B, H, W), np.int32)
    lib().mvfo_grid_sample(ip, gp, _p(out), _p(x0), _p(y0), B, C, H, W)
    return out, x0, y0


def ssim(x, y):
    """layers.py:277-290"""
    x, xp = _f(x)
    y, yp = _f(y)
    B, C, H, W = x.shape
    out = np.empty_like(x)
    lib().mvfo_ssim(xp, yp, _p(out), B * C, H, W)
    return out


def smooth_loss(disp, img):
    """layers.py:231-242"""
    disp, dp = _f(disp)
    img, ip = _f(img)
    B, _, H, W = disp.shape
    return float(lib().mvfo_smooth_loss(dp, ip, B, H, W))


def si_log_loss(pred, target, mask=None, beta=0.5):
    """train.py:924-941"""
    pred, pp = _f(pred)
    target, tp = _f(target)
    mp = None
    if mask is not None:
        mask, mp = _f(mask)
    B, _, H, W = pred.shape
    return float(lib().mvfo_si_log_loss(pp, tp, mp, B, H, W, ctypes.c_float(beta)))


def si_log_loss_bwd(pred, target, mask=None, beta=0.5, gout=1.0):
    pred, pp = _f(pred)
    target, tp = _f(target)
    mp = None
    if mask is not None:
        mask, mp = _f(mask)
    B, _, H, W = pred.shape
    gp = np.empty_like(pred)
    gt = np.empty_like(pred)
    lib().mvfo_si_log_loss_bwd(pp, tp, mp, B, H, W, ctypes.c_float(beta), ctypes.c_float(gout), _p(gp), _p(gt))
    return gp, gt


def f1_forward(disp, tgt, src0, src1, inv_K, P0, P1, noise=None, mask_rec=None, min_depth=0.1,
               max_depth=100.0, smooth_w=1e-3, flags=0, full=True):
    """generate_images_pred x2 + compute_losses_base (train.py:956-1051).  Returns a dict."""
    disp, dp = _f(disp)
    tgt, tp = _f(tgt)
    src0, s0p = _f(src0)
    src1, s1p = _f(src1)
    inv_K, kp = _f(inv_K)
    P0, p0p = _f(P0)
    P1, p1p = _f(P1)
    np_ = mp = None
    if noise is not None:
        noise, np_ = _f(noise)
    if mask_rec is not None:
        mask_rec, mp = _f(mask_rec)
    B, _, H, W = disp.shape
    prm = make_params(B, H, W, min_depth, max_depth, smooth_w, flags)
    loss = np.zeros(3, np.float64)
    to_opt = np.empty((B, H, W), np.float32)
    idx = np.empty((B, H, W), np.uint8)
    out = {"loss": loss, "to_optimise": to_opt, "idx": idx}
    g0 = g1 = w0 = w1 = xy = None
    if full:
        g0 = np.empty((B, H, W, 2), np.float32)
        g1 = np.empty((B, H, W, 2), np.float32)
        w0 = np.empty((B, 3, H, W), np.float32)
        w1 = np.empty((B, 3, H, W), np.float32)
        xy = np.empty((2, 2, B, H, W), np.int32)
        out.update(grid0=g0, grid1=g1, warp0=w0, warp1=w1, x0y0=xy)
    lib().mvfo_f1_forward(ctypes.byref(prm), dp, tp, s0p, s1p, kp, p0p, p1p, np_, mp, _p(loss), _p(to_opt),
                          _p(idx), _p(g0), _p(g1), _p(w0), _p(w1), _p(xy))
    return out


def f1_backward(disp, tgt, src0, src1, inv_K, P0, P1, idx, mask_rec=None, gout=1.0, min_depth=0.1,
                max_depth=100.0, smooth_w=1e-3, flags=0):
    """Analytic gradient of f1_forward w.r.t. disp and P0/P1."""
    disp, dp = _f(disp)
    tgt, tp = _f(tgt)
    src0, s0p = _f(src0)
    src1, s1p = _f(src1)
    inv_K, kp = _f(inv_K)
    P0, p0p = _f(P0)
    P1, p1p = _f(P1)
    mp = None
    if mask_rec is not None:
        mask_rec, mp = _f(mask_rec)
    idx = np.ascontiguousarray(idx, dtype=np.uint8)
    B, _, H, W = disp.shape
    prm = make_params(B, H, W, min_depth, max_depth, smooth_w, flags)
    gd = np.empty_like(disp)
    gP0 = np.empty((B, 3, 4), np.float32)
    gP1 = np.empty((B, 3, 4), np.float32)
    lib().mvfo_f1_backward(ctypes.byref(prm), dp, tp, s0p, s1p, kp, p0p, p1p, mp, _p(idx),
                           ctypes.c_float(gout), _p(gd), _p(gP0), _p(gP1))
    return gd, gP0, gP1
