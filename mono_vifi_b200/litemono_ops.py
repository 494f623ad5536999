"""Host side of csrc/litemono.cu: the HBM-bound pieces of the Lite-Mono blocks (networks/LiteMono.py) as autograd functions over
the C ABI -- depth-wise dilated 3x3 convolution (mvf_dwconv3x3_*), exact GELU (mvf_gelu_*), channels-last LayerNorm
(mvf_layernorm_cl_*).  CUDA tensors only; the network classes route CPU tensors to the torch expressions themselves."""
import torch

from . import _lib

launches = {"dwconv_fwd": 0, "dwconv_dgrad": 0, "dwconv_wgrad": 0, "gelu_fwd": 0, "gelu_bwd": 0, "ln_fwd": 0, "ln_bwd": 0}
_ws = {}


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _workspace(dev, nfloats):
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    ws = _ws.get(key)
    if ws is None or ws.numel() < nfloats:
        ws = torch.empty(max(nfloats, 1), device=dev, dtype=torch.float32)
        _ws[key] = ws
    return ws


def _cl(t):
    """dense channels-last fp32 (NCHW-shaped): element (b,c,y,x) at ((b*H + y)*W + x)*C + c"""
    B, C, H, W = t.shape
    if t.dtype == torch.float32 and tuple(t.stride()) == (H * W * C, 1, W * C, C):
        return t
    out = torch.empty(B, H, W, C, device=t.device, dtype=torch.float32).permute(0, 3, 1, 2)
    out.copy_(t)
    return out


def _empty_cl(B, C, H, W, dev):
    return torch.empty(B, H, W, C, device=dev, dtype=torch.float32).permute(0, 3, 1, 2)


class _DwConv3x3(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, dil):
        x = _cl(x)
        B, C, H, W = x.shape
        w_t = weight.detach().reshape(C, 9).t().contiguous()   # [9][C]
        y = _empty_cl(B, C, H, W, x.device)
        b = None if bias is None else bias.detach().float().contiguous()
        _lib.check(_lib.lib().mvf_dwconv3x3_fwd(x.data_ptr(), w_t.data_ptr(), None if b is None else b.data_ptr(), y.data_ptr(), B, C, H, W, dil, 0,
                                                _stream(x)), "mvf_dwconv3x3_fwd")
        launches["dwconv_fwd"] += 1
        ctx.save_for_backward(x, w_t)
        ctx.dil, ctx.has_bias = dil, bias is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w_t = ctx.saved_tensors
        B, C, H, W = x.shape
        gy = _cl(gy)
        L = _lib.lib()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = _empty_cl(B, C, H, W, x.device)
            _lib.check(L.mvf_dwconv3x3_fwd(gy.data_ptr(), w_t.data_ptr(), None, gx.data_ptr(), B, C, H, W, ctx.dil, 1, _stream(x)), "mvf_dwconv3x3_fwd (dgrad)")
            launches["dwconv_dgrad"] += 1
        if ctx.needs_input_grad[1]:
            gw = torch.empty(C, 1, 3, 3, device=x.device, dtype=torch.float32)
            ws = _workspace(x.device, L.mvf_dwconv3x3_wgrad_workspace_floats(B * H * W, C))
            _lib.check(L.mvf_dwconv3x3_wgrad(x.data_ptr(), gy.data_ptr(), gw.data_ptr(), ws.data_ptr(), ws.numel(), B, C, H, W, ctx.dil, _stream(x)),
                       "mvf_dwconv3x3_wgrad")
            launches["dwconv_wgrad"] += 1
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = gy.sum((0, 2, 3))
        return gx, gw, gb, None


def dwconv3x3_usable(x, weight, stride, padding, dilation, groups):
    C = x.shape[1]
    pair = lambda v: (v, v) if isinstance(v, int) else tuple(v)
    d = pair(dilation)
    return (x.is_cuda and x.dim() == 4 and groups == C and tuple(weight.shape) == (C, 1, 3, 3) and C % 4 == 0 and C <= 1024 and
            pair(stride) == (1, 1) and d[0] == d[1] and pair(padding) == d)


def dwconv3x3(x, weight, bias, dilation):
    """depth-wise 3x3, stride 1, padding = dilation (LiteMono.py:140-155)"""
    d = dilation if isinstance(dilation, int) else dilation[0]
    return _DwConv3x3.apply(x, weight, bias, int(d))


class _Gelu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        if x.dtype != torch.float32:
            x = x.float()
        if not (x.is_contiguous() or (x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last))):
            x = x.contiguous()
        y = torch.empty_like(x)   # preserves the (dense) memory format
        _lib.check(_lib.lib().mvf_gelu_fwd(x.data_ptr(), y.data_ptr(), x.numel(), _stream(x)), "mvf_gelu_fwd")
        launches["gelu_fwd"] += 1
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, gy):
        (x,) = ctx.saved_tensors
        if gy.dtype != torch.float32 or gy.stride() != x.stride():
            g = torch.empty_like(x)
            g.copy_(gy)
            gy = g
        gx = torch.empty_like(x)
        _lib.check(_lib.lib().mvf_gelu_bwd(x.data_ptr(), gy.data_ptr(), gx.data_ptr(), x.numel(), _stream(x)), "mvf_gelu_bwd")
        launches["gelu_bwd"] += 1
        return gx


def gelu(x):
    """nn.GELU() (exact erf form)"""
    if not x.is_cuda or x.numel() % 4:
        return torch.nn.functional.gelu(x)
    return _Gelu.apply(x)


class _LayerNormCL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        C = x.shape[-1]
        shape = x.shape
        x2 = x.reshape(-1, C)
        if x2.dtype != torch.float32 or not x2.is_contiguous():
            x2 = x2.float().contiguous()
        P = x2.shape[0]
        y = torch.empty_like(x2)
        mean = torch.empty(P, device=x.device, dtype=torch.float32)
        rstd = torch.empty(P, device=x.device, dtype=torch.float32)
        w, b = weight.detach().float().contiguous(), bias.detach().float().contiguous()
        _lib.check(_lib.lib().mvf_layernorm_cl_fwd(x2.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), P, C,
                                                   float(eps), _stream(x)), "mvf_layernorm_cl_fwd")
        launches["ln_fwd"] += 1
        ctx.save_for_backward(x2, w, mean, rstd)
        ctx.shape = shape
        return y.view(shape)

    @staticmethod
    def backward(ctx, gy):
        x2, w, mean, rstd = ctx.saved_tensors
        P, C = x2.shape
        gy = gy.reshape(P, C)
        if gy.dtype != torch.float32 or not gy.is_contiguous():
            gy = gy.float().contiguous()
        gx = torch.empty_like(x2)
        gw, gb = torch.empty(C, device=x2.device, dtype=torch.float32), torch.empty(C, device=x2.device, dtype=torch.float32)
        L = _lib.lib()
        ws = _workspace(x2.device, L.mvf_layernorm_bwd_workspace_floats(P, C))
        _lib.check(L.mvf_layernorm_cl_bwd(x2.data_ptr(), gy.data_ptr(), w.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gx.data_ptr(), gw.data_ptr(),
                                          gb.data_ptr(), ws.data_ptr(), ws.numel(), P, C, _stream(x2)), "mvf_layernorm_cl_bwd")
        launches["ln_bwd"] += 1
        return gx.view(ctx.shape), gw, gb, None


def layer_norm_cl(x, weight, bias, eps):
    """F.layer_norm(x, (C,), weight, bias, eps) for tokens with the channels last"""
    C = x.shape[-1]
    if not x.is_cuda or C % 4 or C > 512:
        return torch.nn.functional.layer_norm(x, (C,), weight, bias, eps)
    return _LayerNormCL.apply(x, weight, bias, eps)
