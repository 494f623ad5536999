"""Host side of csrc/warp_cl.cu: feature warp by a flow (IFRNet.warp), bilinear resize (F.interpolate), the PReLU tail and
the pose-matrix kernel, as autograd functions over the C ABI (mvf_flow_warp_*, mvf_resize_bilinear_*, mvf_prelu_cl_fwd,
mvf_pose_matrix_*).  CUDA tensors only go through the library (no fallback there: a missing library raises); CPU tensors
-- the `-m "not gpu"` architecture / state_dict parity tests of the network classes -- take the torch expressions the
reference itself uses."""
import numpy as np
import torch
import torch.nn.functional as F

from . import _lib

launches = {"flow_warp_fwd": 0, "flow_warp_bwd": 0, "resize_fwd": 0, "resize_bwd": 0, "prelu": 0, "pose_matrix": 0}
_ws = {}


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _is_cl(t):
    B, C, H, W = t.shape
    return C % 2 == 0 and tuple(t.stride()) == (H * W * C, 1, W * C, C)


def _layout_of(t):
    """(tensor, layout): dense channels-last with C % 2 == 0 -> 1, otherwise a contiguous NCHW copy -> 0"""
    if t.dtype != torch.float32:
        t = t.float()
    if _is_cl(t):
        return t, 1
    return t.contiguous(), 0


def _empty_like_layout(B, C, H, W, layout, dev):
    if layout == 1:
        return torch.empty(B, H, W, C, device=dev, dtype=torch.float32).permute(0, 3, 1, 2)
    return torch.empty(B, C, H, W, device=dev, dtype=torch.float32)


def _workspace(dev, nbytes):
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    ws = _ws.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        _ws[key] = ws
    return ws


# ---- flow warp -------------------------------------------------------------------------------------------------------
def _warp_torch(img, flow):
    B, _, H, W = flow.shape
    xx = torch.linspace(-1.0, 1.0, W, device=flow.device, dtype=flow.dtype).view(1, 1, 1, W).expand(B, -1, H, -1)
    yy = torch.linspace(-1.0, 1.0, H, device=flow.device, dtype=flow.dtype).view(1, 1, H, 1).expand(B, -1, -1, W)
    grid = torch.cat([xx + flow[:, 0:1] / ((W - 1.0) / 2.0), yy + flow[:, 1:2] / ((H - 1.0) / 2.0)], 1).to(img)
    return F.grid_sample(img, grid.permute(0, 2, 3, 1), mode="bilinear", padding_mode="border", align_corners=True)


class _FlowWarp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, flow):
        x, layout = _layout_of(x)
        flow = flow.contiguous().float()
        B, C, H, W = x.shape
        y = _empty_like_layout(B, C, H, W, layout, x.device)
        _lib.check(_lib.lib().mvf_flow_warp_fwd(x.data_ptr(), flow.data_ptr(), y.data_ptr(), B, C, H, W, layout, _stream(x)), "mvf_flow_warp_fwd")
        launches["flow_warp_fwd"] += 1
        ctx.save_for_backward(flow)
        ctx.dims = (B, C, H, W)
        return y

    @staticmethod
    def backward(ctx, gy):
        (flow,) = ctx.saved_tensors
        B, C, H, W = ctx.dims
        if C % 2:
            raise RuntimeError("flow_warp backward needs an even channel count (feature maps); image warps carry no gradient on this path")
        if not _is_cl(gy) or gy.dtype != torch.float32:
            g = torch.empty(B, H, W, C, device=gy.device, dtype=torch.float32).permute(0, 3, 1, 2)
            g.copy_(gy)
            gy = g
        gx = torch.empty(B, H, W, C, device=gy.device, dtype=torch.float32).permute(0, 3, 1, 2)
        L = _lib.lib()
        nbytes = L.mvf_flow_warp_bwd_workspace_bytes(B, C, H, W)
        ws = _workspace(gy.device, nbytes)
        _lib.check(L.mvf_flow_warp_bwd(gy.data_ptr(), flow.data_ptr(), gx.data_ptr(), B, C, H, W, ws.data_ptr(), ws.numel(), _stream(gy)),
                   "mvf_flow_warp_bwd")
        launches["flow_warp_bwd"] += 1
        return gx, None


def flow_warp(img, flow):
    """IFRNet.py:7-15: backward warp of img [B,C,H,W] by the pixel-unit flow [B,2,H,W]; border padding, align_corners=True.
    The flow is treated as a constant (it comes from the frozen VFI network, train.py:210-216)."""
    if not img.is_cuda:
        return _warp_torch(img, flow)
    if flow.requires_grad:
        raise RuntimeError("flow_warp: gradients w.r.t. the flow are not on the hot path (the VFI network is frozen)")
    if img.requires_grad and img.shape[1] % 2:
        raise RuntimeError("flow_warp: a differentiable warp needs an even channel count")
    return _FlowWarp.apply(img, flow.detach())


# ---- bilinear resize -------------------------------------------------------------------------------------------------
def _scales(in_size, out_size, scale_factor, align_corners):
    """torch's area_pixel_compute_scale for float tensors"""
    if align_corners:
        return float(np.float32(in_size - 1) / np.float32(out_size - 1)) if out_size > 1 else 0.0
    if scale_factor is not None:
        return float(np.float32(1.0 / scale_factor))
    return float(np.float32(in_size) / np.float32(out_size))


class _Resize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, Ho, Wo, sh, sw, align, mul):
        x, layout = _layout_of(x)
        B, C, Hi, Wi = x.shape
        if layout == 1 and mul != (1.0, 1.0):
            x, layout = x.contiguous(), 0
        y = _empty_like_layout(B, C, Ho, Wo, layout, x.device)
        _lib.check(_lib.lib().mvf_resize_bilinear_fwd(x.data_ptr(), y.data_ptr(), B, C, Hi, Wi, Ho, Wo, sh, sw, int(align), mul[0], mul[1],
                                                      layout, _stream(x)), "mvf_resize_bilinear_fwd")
        launches["resize_fwd"] += 1
        ctx.args = (B, C, Hi, Wi, Ho, Wo, sh, sw, int(align), mul, layout)
        return y

    @staticmethod
    def backward(ctx, gy):
        B, C, Hi, Wi, Ho, Wo, sh, sw, align, mul, layout = ctx.args
        if layout == 1:
            if not _is_cl(gy) or gy.dtype != torch.float32:
                g = torch.empty(B, Ho, Wo, C, device=gy.device, dtype=torch.float32).permute(0, 3, 1, 2)
                g.copy_(gy)
                gy = g
        else:
            gy = gy.contiguous().float()
        gx = _empty_like_layout(B, C, Hi, Wi, layout, gy.device)
        _lib.check(_lib.lib().mvf_resize_bilinear_bwd(gy.data_ptr(), gx.data_ptr(), B, C, Hi, Wi, Ho, Wo, sh, sw, align, mul[0], mul[1], layout,
                                                      _stream(gy)), "mvf_resize_bilinear_bwd")
        launches["resize_bwd"] += 1
        return gx, None, None, None, None, None, None


def resize_bilinear(x, size=None, scale_factor=None, align_corners=False, mul=(1.0, 1.0)):
    """F.interpolate(x, size | scale_factor, mode="bilinear", align_corners=...) [* per-channel multipliers (even, odd)]."""
    Hi, Wi = x.shape[-2:]
    if size is not None:
        Ho, Wo = int(size[0]), int(size[1])
        sfh = sfw = None
    else:
        sfh, sfw = (scale_factor, scale_factor) if not isinstance(scale_factor, (tuple, list)) else scale_factor
        Ho, Wo = int(np.floor(Hi * sfh)), int(np.floor(Wi * sfw))
    if not x.is_cuda:
        y = F.interpolate(x, size=size, scale_factor=scale_factor, mode="bilinear", align_corners=align_corners)
        if mul != (1.0, 1.0):
            y = y * y.new_tensor([mul[i & 1] for i in range(y.shape[1])]).view(1, -1, 1, 1)
        return y
    sh, sw = _scales(Hi, Ho, sfh, align_corners), _scales(Wi, Wo, sfw, align_corners)
    return _Resize.apply(x, Ho, Wo, sh, sw, bool(align_corners), (float(mul[0]), float(mul[1])))


# ---- PReLU tail (inference) ------------------------------------------------------------------------------------------
def prelu(x, slope, res=None):
    """nn.PReLU(C)(x + res) on a channels-last tensor (forward only: the VFI network is frozen)."""
    if not x.is_cuda or x.shape[1] % 4 or torch.is_grad_enabled() and (x.requires_grad or slope.requires_grad):
        return F.prelu(x if res is None else x + res, slope)
    x, layout = _layout_of(x)
    if layout != 1:
        x = x.contiguous(memory_format=torch.channels_last)
    if res is not None:
        res = res.float()
        if not _is_cl(res):
            res = res.contiguous(memory_format=torch.channels_last)
    B, C, H, W = x.shape
    y = torch.empty(B, H, W, C, device=x.device, dtype=torch.float32).permute(0, 3, 1, 2)
    _lib.check(_lib.lib().mvf_prelu_cl_fwd(x.data_ptr(), None if res is None else res.data_ptr(), slope.detach().contiguous().data_ptr(),
                                           y.data_ptr(), B * H * W, C, _stream(x)), "mvf_prelu_cl_fwd")
    launches["prelu"] += 1
    return y


# ---- pose matrix -----------------------------------------------------------------------------------------------------
class _PoseMatrix(torch.autograd.Function):
    @staticmethod
    def forward(ctx, axisangle, translation, invert):
        aa = axisangle.reshape(-1, 3).contiguous().float()
        tr = translation.reshape(-1, 3).contiguous().float()
        B = aa.shape[0]
        M = torch.empty(B, 4, 4, device=aa.device, dtype=torch.float32)
        _lib.check(_lib.lib().mvf_pose_matrix_fwd(aa.data_ptr(), tr.data_ptr(), M.data_ptr(), B, int(invert), _stream(aa)), "mvf_pose_matrix_fwd")
        launches["pose_matrix"] += 1
        ctx.save_for_backward(aa, tr)
        ctx.invert, ctx.shapes = int(invert), (axisangle.shape, translation.shape)
        return M

    @staticmethod
    def backward(ctx, gM):
        aa, tr = ctx.saved_tensors
        B = aa.shape[0]
        gM = gM.contiguous().float()
        gaa, gtr = torch.empty_like(aa), torch.empty_like(tr)
        _lib.check(_lib.lib().mvf_pose_matrix_bwd(aa.data_ptr(), tr.data_ptr(), gM.data_ptr(), gaa.data_ptr(), gtr.data_ptr(), B, ctx.invert,
                                                  _stream(aa)), "mvf_pose_matrix_bwd")
        return gaa.view(ctx.shapes[0]), gtr.view(ctx.shapes[1]), None


def pose_matrix(axisangle, translation, invert=False):
    """transformation_from_parameters (layers.py:28-45) as one kernel: [B,1,3] x 2 -> [B,4,4]"""
    return _PoseMatrix.apply(axisangle, translation, invert)
