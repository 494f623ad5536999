"""Fused data movement of the depth decoder on channels-last tensors: nearest-upsample x2 + channel concat +
ReflectionPad2d(1) in one kernel (forward and exact adjoint), instead of the reference's F.interpolate + torch.cat +
nn.ReflectionPad2d (monodepth2.py:86-93, layers.py:126-139).  C ABI: mvf_upcat_pad_fwd / _bwd."""
import torch

from . import _lib


def _dense_cl(t):
    """NCHW-shaped, dense channels-last: element (b,c,y,x) at ((b*H + y)*W + x)*C + c."""
    B, C, H, W = t.shape
    want = (H * W * C, 1, W * C, C)
    if t.dtype != torch.float32:
        t = t.float()
    if tuple(t.stride()) != want:
        out = torch.empty(B, H, W, C, device=t.device, dtype=torch.float32).permute(0, 3, 1, 2)
        out.copy_(t)
        t = out
    return t


def usable(a, skip=None, upsample=False):
    if not a.is_cuda or a.dim() != 4 or a.shape[1] % 4:
        return False
    H, W = (a.shape[2] * 2, a.shape[3] * 2) if upsample else (a.shape[2], a.shape[3])
    if H < 4 or W < 4:
        return False
    if skip is not None and (skip.shape[1] % 4 or tuple(skip.shape[2:]) != (H, W) or skip.shape[0] != a.shape[0]):
        return False
    return True


class _UpcatPad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, skip, upsample):
        a = _dense_cl(a)
        B, Ca, Ha, Wa = a.shape
        H, W = (2 * Ha, 2 * Wa) if upsample else (Ha, Wa)
        Cs = 0
        if skip is not None:
            skip = _dense_cl(skip)
            Cs = skip.shape[1]
        y = torch.empty(B, H + 2, W + 2, Ca + Cs, device=a.device, dtype=torch.float32).permute(0, 3, 1, 2)
        st = torch.cuda.current_stream(a.device).cuda_stream
        _lib.check(_lib.lib().mvf_upcat_pad_fwd(a.data_ptr(), None if skip is None else skip.data_ptr(), y.data_ptr(), B, Ca, Cs,
                                                H, W, 1 if upsample else 0, st), "mvf_upcat_pad_fwd")
        ctx.dims = (B, Ca, Cs, H, W, upsample, Ha, Wa)
        return y

    @staticmethod
    def backward(ctx, gy):
        B, Ca, Cs, H, W, upsample, Ha, Wa = ctx.dims
        gy = _dense_cl(gy)
        ga = gs = None
        if ctx.needs_input_grad[0]:
            ga = torch.empty(B, Ha, Wa, Ca, device=gy.device, dtype=torch.float32).permute(0, 3, 1, 2)
        if Cs and ctx.needs_input_grad[1]:
            gs = torch.empty(B, H, W, Cs, device=gy.device, dtype=torch.float32).permute(0, 3, 1, 2)
        st = torch.cuda.current_stream(gy.device).cuda_stream
        _lib.check(_lib.lib().mvf_upcat_pad_bwd(gy.data_ptr(), None if ga is None else ga.data_ptr(),
                                                None if gs is None else gs.data_ptr(), B, Ca, Cs, H, W, 1 if upsample else 0, st),
                   "mvf_upcat_pad_bwd")
        return ga, gs, None


def upcat_pad(a, skip=None, upsample=False):
    """ReflectionPad2d(1)(cat([upsample(a) if upsample else a, skip], 1)) -> [B, Ca+Cs, H+2, W+2], channels-last."""
    return _UpcatPad.apply(a, skip, bool(upsample))


class _MaxPool3s2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = _dense_cl(x)
        B, C, H, W = x.shape
        Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        y = torch.empty(B, Ho, Wo, C, device=x.device, dtype=torch.float32).permute(0, 3, 1, 2)
        idx = torch.empty(B, Ho, Wo, C, device=x.device, dtype=torch.uint8)
        st = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(_lib.lib().mvf_maxpool3s2_fwd(x.data_ptr(), y.data_ptr(), idx.data_ptr(), B, C, H, W, st), "mvf_maxpool3s2_fwd")
        ctx.save_for_backward(idx)
        ctx.dims = (B, C, H, W)
        return y

    @staticmethod
    def backward(ctx, gy):
        (idx,) = ctx.saved_tensors
        B, C, H, W = ctx.dims
        gy = _dense_cl(gy)
        gx = torch.empty(B, H, W, C, device=gy.device, dtype=torch.float32).permute(0, 3, 1, 2)
        st = torch.cuda.current_stream(gy.device).cuda_stream
        _lib.check(_lib.lib().mvf_maxpool3s2_bwd(gy.data_ptr(), idx.data_ptr(), gx.data_ptr(), B, C, H, W, st), "mvf_maxpool3s2_bwd")
        return gx


def maxpool3s2(x):
    """nn.MaxPool2d(kernel_size=3, stride=2, padding=1) on a channels-last CUDA tensor (C % 4 == 0)."""
    return _MaxPool3s2.apply(x)


class MaxPool3s2(torch.nn.Module):
    """parameter-free drop-in for torchvision ResNet's `maxpool` (same position in the module tree, no state_dict keys)"""

    def forward(self, x):
        if x.is_cuda and x.dim() == 4 and x.shape[1] % 4 == 0:
            return maxpool3s2(x)
        return torch.nn.functional.max_pool2d(x, 3, 2, 1)
