"""Tensor-core convolutions (tcgen05 / TMEM / TMA implicit GEMM on NCHW fp32, TF32 inputs, fp32 accumulate).

`conv2d(x, weight, bias, stride, padding)` is a torch.autograd.Function over the C ABI in
include/monovifi_b200.h (mvf_conv2d_*).  Forward and the input gradient run the same kernel (the latter on
grad_out with the flipped / transposed filter bank); the weight gradient has its own kernel.  `supported()`
tells the dispatcher in conv.py which problems the kernels cover; nothing here falls back silently.
Activations are NCHW-shaped tensors in torch.channels_last memory format (the kernels' TMA boxes need channels
contiguous); tensors arriving in another layout are converted once on entry.
"""
import os
import weakref

import torch
import torch.nn.functional as F

from . import _lib

launches = {"fprop": 0, "dgrad": 0, "wgrad": 0, "pack": 0}
# when `timing` is a list, every tensor-core launch is bracketed by CUDA events on the launching stream and recorded
# as (tag, algorithmic flops, start, end) -- bench.py derives the achieved TFLOP/s of the conv kernels from it
timing = None
_tag = "fprop"


shapes = None   # a list: the (B, Cin, H, W, Cout, KH, KW, stride) of every timed launch, parallel to `timing` (tools/conv_layers.py)


def _timed(tag, flops, fn, shape=None):
    if timing is None:
        return fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = fn()
    e1.record()
    timing.append((tag, flops, e0, e1))
    if shapes is not None:
        shapes.append(shape)
    return r


def _pair(v):
    return (v, v) if isinstance(v, int) else (int(v[0]), int(v[1]))


def _desc(B, Cin, H, W, Cout, KH, KW, pad, stride, x, y):
    sy, sx = _pair(stride)
    d = _lib.Conv2dDesc(B, Cin, H, W, Cout, KH, KW, pad, sy)
    d.stride_x = sx
    for i, dim in enumerate((0, 2, 3)):  # batch, row, column strides; the channel stride is 1
        d.x_stride[i] = x.stride(dim)
        d.y_stride[i] = y.stride(dim)
    return d


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _as_input(x):
    """NCHW-shaped fp32 tensor in channels-last memory (unit channel stride, 16-byte aligned pixels): what the
    TMA descriptor of the kernels needs.  Channel slices of a wider channels-last tensor qualify as they are."""
    if x.dtype != torch.float32:
        x = x.float()
    sB, sC, sH, sW = x.stride()
    if (sC != 1 and x.shape[1] != 1) or sW % 4 or sH % 4 or sB % 4 or x.data_ptr() % 16:
        x = x.contiguous(memory_format=torch.channels_last)
        if x.stride(1) != 1:  # torch keeps NCHW strides for some degenerate shapes (C == 1 or H == W == 1)
            B, C, H, W = x.shape
            x = x.as_strided((B, C, H, W), (H * W * C, 1, W * C, C))
    return x


def out_hw(H, W, KH, KW, pad, stride):
    sy, sx = _pair(stride)
    return (H + 2 * pad - KH) // sy + 1, (W + 2 * pad - KW) // sx + 1


def supported(x, weight, stride, padding, dilation=1, groups=1):
    def pair(v):
        return (v, v) if isinstance(v, int) else tuple(v)
    if not x.is_cuda or x.dim() != 4 or groups != 1 or pair(dilation) != (1, 1):
        return False
    sh, sw = pair(stride)
    ph, pw = pair(padding)
    Cout, Cin, KH, KW = weight.shape
    if sh not in (1, 2) or sw not in (1, 2) or ph != pw or Cin % 4:
        return False
    return True


# Packed banks of PARAMETERS are reused while the weights are unchanged: a network applied twice per step (the two pose
# passes) packs once.  "Unchanged" = same storage, same tensor version, same optimiser epoch (FlatAdamW updates the
# arena through raw pointers and bumps `weights_epoch`), same stream and same graph capture (a recorded graph must
# contain its own pack launches).
weights_epoch = 0
_pack_cache = {}


class FilterBank:
    """All filter banks of a training step packed by ONE kernel launch (mvf_conv2d_pack_filters_multi) at the start of the step,
    instead of one launch per layer and direction inside the forward / backward (155 per step of the ResNet18 configuration).
    `weights`: 4-D convolution weights and 2-D nn.Linear weights (used as 1x1 convolutions) that change every step;
    `frozen`: weights of networks that are never trained (the VFI network) -- packed once, re-packed only if their version changes.
    pack_filters() consults the active bank first; anything not registered (zero-padded or reshaped temporaries) is packed per call."""

    def __init__(self, weights, frozen=()):
        self._args = (list(weights), list(frozen))
        self._build()

    def _build(self):
        """(re)build the pointer tables: the flat-arena optimiser re-points every parameter's storage at its first step"""
        L = _lib.lib()
        weights, frozen = self._args
        self.chunk = L.mvf_conv2d_pack_chunk()
        self.sets = []
        self._ptrs = [w.data_ptr() for w in weights + frozen]
        for ws, is_frozen in ((list(weights), False), (list(frozen), True)):
            ws = [w for w in ws if w.is_cuda and w.dim() in (2, 4) and w.dtype == torch.float32 and w.is_contiguous()]
            if not ws:
                continue
            shapes = [tuple(w.shape) if w.dim() == 4 else (w.shape[0], w.shape[1], 1, 1) for w in ws]
            sizes = []
            for (Cout, Cin, KH, KW) in shapes:
                for dg in (0, 1):
                    N, K = (Cin, Cout) if dg else (Cout, Cin)
                    sizes.append(L.mvf_conv2d_packed_filter_floats(N, K, KH, KW))
            arena = torch.empty(sum((n + 3) // 4 * 4 for n in sizes), device=ws[0].device, dtype=torch.float32)
            rows, views, off, blk, j = [], {}, 0, 0, 0
            for wi, (w, (Cout, Cin, KH, KW)) in enumerate(zip(ws, shapes)):
                for dg in (0, 1):
                    n = sizes[j]
                    j += 1
                    view = arena[off:off + n]
                    rows.append([w.data_ptr(), view.data_ptr(), Cout, Cin, KH, KW, dg, blk])
                    views[(w.data_ptr(), Cout, Cin, KH, KW, bool(dg))] = (view, wi)
                    off += (n + 3) // 4 * 4
                    blk += (n + self.chunk - 1) // self.chunk
            self.sets.append({"frozen": is_frozen, "weights": ws, "arena": arena, "views": views, "blocks": blk, "n": len(rows),
                              "table": torch.tensor(rows, dtype=torch.int64).to(ws[0].device), "tag": None, "versions": None})

    def refresh(self):
        """pack everything that may have changed; call at the start of a step, on the stream the step starts on"""
        global _bank
        L = _lib.lib()
        if self._ptrs != [w.data_ptr() for w in self._args[0] + self._args[1]]:
            self._build()
        for st in self.sets:
            dev = st["arena"].device
            stream = torch.cuda.current_stream(dev).cuda_stream
            if st["frozen"]:
                versions = [w._version for w in st["weights"]]
                if st["versions"] == versions:
                    continue
                st["versions"] = versions
            _lib.check(L.mvf_conv2d_pack_filters_multi(st["table"].data_ptr(), st["n"], st["blocks"], stream), "mvf_conv2d_pack_filters_multi")
            launches["pack"] += 1
            st["tag"] = (weights_epoch, L.mvf_stream_capture_id(stream), [w._version for w in st["weights"]])
        _bank = self

    def lookup(self, weight, dgrad):
        Cout, Cin, KH, KW = weight.shape
        key = (weight.data_ptr(), Cout, Cin, KH, KW, bool(dgrad))
        for st in self.sets:
            hit = st["views"].get(key)
            if hit is None:
                continue
            view, wi = hit
            tag = st["tag"]
            if tag is None or tag[2][wi] != weight._version:   # (a view of a parameter shares its version counter)
                return None
            if st["frozen"]:
                return view
            if tag[0] == weights_epoch and tag[1] == _lib.lib().mvf_stream_capture_id(_stream(weight)):
                return view
            return None
        return None

    def release(self):
        global _bank
        if _bank is self:
            _bank = None


_bank = None


def pack_filters(weight, dgrad=False):
    if _bank is not None and weight.is_cuda:
        hit = _bank.lookup(weight, dgrad)
        if hit is not None:
            return hit
    Cout, Cin, KH, KW = weight.shape
    N, K = (Cin, Cout) if dgrad else (Cout, Cin)
    key = tag = None
    if weight.requires_grad and weight.is_leaf and weight.is_cuda:
        st = _stream(weight)
        key = (weight.data_ptr(), tuple(weight.shape), dgrad, st)
        tag = (weight._version, weights_epoch, _lib.lib().mvf_stream_capture_id(st))
        hit = _pack_cache.get(key)
        if hit is not None and hit[0] == tag and hit[2]() is weight:  # same Parameter object, not a recycled address
            return hit[1]
    n = _lib.lib().mvf_conv2d_packed_filter_floats(N, K, KH, KW)
    out = torch.empty(n, device=weight.device, dtype=torch.float32)
    if key is not None:
        _pack_cache[key] = (tag, out, weakref.ref(weight))
    w = weight.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    launches["pack"] += 1
    _lib.check(_lib.lib().mvf_conv2d_pack_filters(w.data_ptr(), out.data_ptr(), Cout, Cin, KH, KW, 1 if dgrad else 0,
                                                  _stream(weight)), "mvf_conv2d_pack_filters")
    return out


def conv_forward_raw(x, w_packed, bias, Cout, KH, KW, pad, stride=1, act=0, out=None):
    """y = conv(x) with a packed filter bank; x and y are NCHW-shaped, channels-last in memory."""
    x = _as_input(x)
    B, Cin, H, W = x.shape
    Ho, Wo = out_hw(H, W, KH, KW, pad, stride)
    if out is None:
        y = torch.empty(B, Ho, Wo, Cout, device=x.device, dtype=torch.float32).permute(0, 3, 1, 2)
    else:
        y = out
        assert (y.stride(1) == 1 or Cout == 1) and tuple(y.shape) == (B, Cout, Ho, Wo)
    d = _desc(B, Cin, H, W, Cout, KH, KW, pad, stride, x, y)
    b = None if bias is None else bias.detach().float().contiguous()
    flops = 2.0 * B * Ho * Wo * Cout * Cin * KH * KW
    rc = _timed(_tag, flops, lambda: _lib.lib().mvf_conv2d_forward(
        d, x.data_ptr(), w_packed.data_ptr(), None if b is None else b.data_ptr(), y.data_ptr(), act, _stream(x)),
        (B, Cin, H, W, Cout, KH, KW, stride))
    _lib.check(rc, "mvf_conv2d_forward")
    return y


ACT = {None: 0, "none": 0, "relu": 1, "elu": 2}


def conv2d_prelu_inference(x, weight, bias, slope, stride=1, padding=0):
    """conv + bias + nn.PReLU(Cout) in the kernel's epilogue (mvf_conv2d_forward_prelu): the `convrelu` block of the frozen VFI
    network (IFRNet.py:121-125).  Forward only."""
    pad = padding if isinstance(padding, int) else padding[0]
    x = _as_input(x)
    B, Cin, H, W = x.shape
    Cout, _, KH, KW = weight.shape
    Ho, Wo = out_hw(H, W, KH, KW, pad, stride)
    y = torch.empty(B, Ho, Wo, Cout, device=x.device, dtype=torch.float32).permute(0, 3, 1, 2)
    d = _desc(B, Cin, H, W, Cout, KH, KW, pad, stride, x, y)
    b = None if bias is None else bias.detach().float().contiguous()
    sl = slope.detach().float().contiguous()
    launches["fprop"] += 1
    flops = 2.0 * B * Ho * Wo * Cout * Cin * KH * KW
    rc = _timed("fprop", flops, lambda: _lib.lib().mvf_conv2d_forward_prelu(
        d, x.data_ptr(), pack_filters(weight).data_ptr(), None if b is None else b.data_ptr(), sl.data_ptr(), y.data_ptr(), _stream(x)))
    _lib.check(rc, "mvf_conv2d_forward_prelu")
    return y


def conv_transpose2d_s2_inference(x, weight, bias, padding):
    """nn.ConvTranspose2d(Cin_t, Cout_t, k, stride 2, padding) forward on the stride-2 data-gradient kernel
    (mvf_conv_transpose2d_s2_fwd; IFRNet.py:194).  weight [Cin_t, Cout_t, k, k].  Channel counts that are not multiples of 4
    are zero-padded (weight, bias, and the input's channels) and the result sliced back.  Forward only."""
    Cin_t, Cout_t, KH, KW = weight.shape
    B, _, H, W = x.shape
    Ho, Wo = (H - 1) * 2 - 2 * padding + KH, (W - 1) * 2 - 2 * padding + KW
    w = weight.detach()
    b = None if bias is None else bias.detach().float()
    ep_out, ep_in = (-Cout_t) % 4, (-Cin_t) % 4
    if ep_out or ep_in:
        w = F.pad(w, (0, 0, 0, 0, 0, ep_out, 0, ep_in))
        b = None if b is None else F.pad(b, (0, ep_out))
        if ep_in:
            x = F.pad(x, (0, 0, 0, 0, 0, ep_in))
    x = _as_input(x)
    Ci, Co = Cin_t + ep_in, Cout_t + ep_out
    y = torch.empty(B, Ho, Wo, Co, device=x.device, dtype=torch.float32).permute(0, 3, 1, 2)
    d = _desc(B, Co, Ho, Wo, Ci, KH, KW, padding, 2, y, x)   # the stride-2 convolution whose data gradient this is
    launches["dgrad"] += 1
    flops = 2.0 * B * H * W * Ci * Co * KH * KW
    wp = pack_filters(w if (ep_out or ep_in) else weight, dgrad=True)
    rc = _timed("dgrad", flops, lambda: _lib.lib().mvf_conv_transpose2d_s2_fwd(
        d, x.data_ptr(), wp.data_ptr(), None if b is None else b.contiguous().data_ptr(), y.data_ptr(), _stream(x)))
    _lib.check(rc, "mvf_conv_transpose2d_s2_fwd")
    return y[:, :Cout_t] if ep_out else y


def _dense_cl(t):
    B, C, H, W = t.shape
    return t.dtype == torch.float32 and tuple(t.stride()) == (H * W * C, 1, W * C, C) and t.data_ptr() % 16 == 0


_bias_ws = {}


def act_bwd_bias(gy, y, act, want_bias):
    """(gy * act'(y), sum over pixels of that) in one pass (C ABI: mvf_act_bwd_bias); dense channels-last inputs"""
    B, C, H, W = gy.shape
    P = B * H * W
    L = _lib.lib()
    gpre = torch.empty(B, H, W, C, device=gy.device, dtype=torch.float32).permute(0, 3, 1, 2) if act else None
    gb = torch.empty(C, device=gy.device, dtype=torch.float32) if want_bias else None
    ws = None
    if want_bias:
        n = L.mvf_bn_workspace_floats(P, C)
        key = (gy.device.index, _stream(gy))
        ws = _bias_ws.get(key)
        if ws is None or ws.numel() < n:
            ws = _bias_ws[key] = torch.empty(max(n, 1 << 18), device=gy.device, dtype=torch.float32)
    _lib.check(L.mvf_act_bwd_bias(gy.data_ptr(), None if y is None else y.data_ptr(), None if gpre is None else gpre.data_ptr(),
                                  None if gb is None else gb.data_ptr(), None if ws is None else ws.data_ptr(),
                                  0 if ws is None else ws.numel(), P, C, act, _stream(gy)), "mvf_act_bwd_bias")
    return (gpre if act else gy), gb


wgrad_stream_enabled = os.environ.get("MVF_WGRAD_STREAM", "1") != "0"
_companions = {}


def _companion(cur):
    key = (cur.device.index, cur.cuda_stream)
    st = _companions.get(key)
    if st is None:
        st = _companions[key] = torch.cuda.Stream(device=cur.device)
    return st


class _Conv2dTC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, pad, stride, act):
        Cout, Cin, KH, KW = weight.shape
        launches["fprop"] += 1
        x = _as_input(x)
        y = conv_forward_raw(x, pack_filters(weight), bias, Cout, KH, KW, pad, stride, act=act)  # activation in the epilogue
        if act:
            ctx.save_for_backward(x, weight, y)
        else:
            ctx.save_for_backward(x, weight)
        ctx.pad, ctx.stride, ctx.has_bias, ctx.act = pad, stride, bias is not None, act
        return y

    @staticmethod
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors[:2]
        Cout, Cin, KH, KW = weight.shape
        pad, stride = ctx.pad, ctx.stride
        gx = gw = gb = None
        want_gb = ctx.has_bias and ctx.needs_input_grad[2]
        yact = ctx.saved_tensors[2] if ctx.act else None
        if (ctx.act or want_gb) and _dense_cl(gy) and (yact is None or _dense_cl(yact)) and Cout % 4 == 0 and Cout <= 1024:
            gy, gb = act_bwd_bias(gy, yact, ctx.act, want_gb)   # one pass: activation backward + bias gradient
            want_gb = False
        elif ctx.act == 1:
            gy = torch.ops.aten.threshold_backward(gy, yact, 0.0)
        elif ctx.act == 2:  # d elu / d pre-activation from the saved OUTPUT: 1 where y > 0, y + 1 elsewhere
            gy = torch.ops.aten.elu_backward(gy, 1.0, 1.0, 1.0, True, yact)
        gy = _as_input(gy)
        # dL/dw and dL/dx of one layer are independent: the weight gradient goes to a companion stream so that the two
        # kernels (each 1-3 tiles per SM) share the GPU; joined before this node returns, so nothing downstream changes
        fork = None
        if ctx.needs_input_grad[1] and ctx.needs_input_grad[0] and wgrad_stream_enabled and timing is None and gy.is_cuda:
            cur = torch.cuda.current_stream(gy.device)
            fork = _companion(cur)
            fork.wait_stream(cur)
            with torch.cuda.stream(fork):
                gw = weight_grad(x, gy, weight.shape, pad, stride)
        if ctx.needs_input_grad[0]:
            if _pair(stride) == (1, 1) and pad <= KH - 1 and pad <= KW - 1:
                gyd, wd = gy, weight
                if Cout % 4:  # e.g. the 1-channel disparity head: zero channels up to a 16-byte pixel
                    extra = 4 - Cout % 4
                    gyd = torch.zeros(gy.shape[0], gy.shape[2], gy.shape[3], Cout + extra, device=gy.device).permute(0, 3, 1, 2)
                    gyd[:, :Cout] = gy
                    wd = torch.nn.functional.pad(weight, (0, 0, 0, 0, 0, 0, 0, extra))
                global _tag
                launches["dgrad"] += 1
                _tag = "dgrad"
                try:
                    gx = conv_forward_raw(gyd, pack_filters(wd, dgrad=True), None, Cin, KH, KW, KH - 1 - pad)
                finally:
                    _tag = "fprop"
            elif dgrad_s2_enabled and _pair(stride) == (2, 2) and Cout % 4 == 0 and Cin % 4 == 0 and KH <= 8 and KW <= 8:
                gx = input_grad_s2(x, gy, weight, pad)
            else:
                gx = input_grad_library(x, gy, weight, pad, stride)
        if fork is not None:
            cur.wait_stream(fork)
        elif ctx.needs_input_grad[1]:
            gw = weight_grad(x, gy, weight.shape, pad, stride)
        if want_gb:
            gb = gy.sum((0, 2, 3))
        return gx, gw, gb, None, None, None


# Stride-2 data gradient on the tcgen05 path (mvf_conv2d_dgrad_s2): class / tap plan checked on the CPU
# (tests/test_dgrad_s2_plan.py), kernel against fp64 torch on the GPU (tests/test_conv_tc_cuda.py).  MVF_DGRAD_S2=0 routes these
# gradients to cuDNN for A/B timing; they are then counted in conv.stats["cudnn_dgrad"].
dgrad_s2_enabled = os.environ.get("MVF_DGRAD_S2", "1") == "1"


def input_grad_s2(x, gy, weight, pad):
    """dL/dx of a stride-2 convolution: four parity classes, each a small stride-1 convolution over gy (one launch)"""
    Cout, Cin, KH, KW = weight.shape
    B, _, H, W = x.shape
    gy = _as_input(gy)
    # classes without taps are not written by the kernel (1 x k / k x 1 filters): start from zeros then
    alloc = torch.zeros if (KH == 1 or KW == 1) else torch.empty
    gx = alloc(B, H, W, Cin, device=gy.device, dtype=torch.float32).permute(0, 3, 1, 2)
    d = _desc(B, Cin, H, W, Cout, KH, KW, pad, 2, gx, gy)
    launches["dgrad"] += 1
    Ho, Wo = gy.shape[2], gy.shape[3]
    flops = 2.0 * B * Ho * Wo * Cout * Cin * KH * KW
    _lib.check(_timed("dgrad", flops, lambda: _lib.lib().mvf_conv2d_dgrad_s2(d, gy.data_ptr(), pack_filters(weight, dgrad=True).data_ptr(),
                                                                            gx.data_ptr(), _stream(gy)), (B, Cin, H, W, Cout, KH, KW, 2)),
               "mvf_conv2d_dgrad_s2")
    return gx


def input_grad_library(x, gy, weight, pad, stride):
    """dL/dx of the shapes the tcgen05 dgrad does not cover yet (strided convolutions): cuDNN, counted in conv.stats."""
    from . import conv
    if os.environ.get("MVF_DGRAD_S2", "1") != "0":   # (MVF_DGRAD_S2=0 is the explicit A/B switch to the library)
        conv.require_fallback("data gradient of conv2d(x=%s, weight=%s, stride=%s, padding=%s)" % (tuple(x.shape), tuple(weight.shape), stride, pad))
    conv.stats["cudnn_dgrad"] = conv.stats.get("cudnn_dgrad", 0) + 1
    gx, _, _ = torch.ops.aten.convolution_backward(gy, x, weight, None, list(_pair(stride)), [pad, pad], [1, 1], False, [0, 0], 1,
                                                   [True, False, False])
    return gx


_wgrad_ws = {}


def weight_grad(x, gy, wshape, pad, stride=1):
    """dL/dw on the tcgen05 wgrad kernel; shapes it does not cover (7x7 stem, odd channel counts) go to cuDNN and
    are counted in conv.stats -- never silently."""
    Cout, Cin, KH, KW = wshape
    B, _, H, W = x.shape
    if Cout % 4:
        # e.g. the 1-channel disparity head: zero channels up to a 16-byte pixel, slice the result back
        extra = 4 - Cout % 4
        gyp = torch.zeros(B, gy.shape[2], gy.shape[3], Cout + extra, device=gy.device, dtype=torch.float32).permute(0, 3, 1, 2)
        gyp[:, :Cout] = gy
        return weight_grad(x, gyp, (Cout + extra, Cin, KH, KW), pad, stride)[:Cout]
    gy = _as_input(gy)
    d = _desc(B, Cin, H, W, Cout, KH, KW, pad, stride, x, gy)
    L = _lib.lib()
    if L.mvf_conv2d_wgrad_supported(d):
        n = L.mvf_conv2d_wgrad_workspace_floats(d)
        key = (x.device.index, _stream(x))
        ws = _wgrad_ws.get(key)
        if ws is None or ws.numel() < n:
            ws = torch.empty(max(n, 1 << 20), device=x.device, dtype=torch.float32)
            _wgrad_ws[key] = ws
        gw = torch.empty(Cout, Cin, KH, KW, device=x.device, dtype=torch.float32)
        launches["wgrad"] += 1
        Ho, Wo = out_hw(H, W, KH, KW, pad, stride)
        flops = 2.0 * B * Ho * Wo * Cout * Cin * KH * KW
        _lib.check(_timed("wgrad", flops, lambda: L.mvf_conv2d_wgrad(d, x.data_ptr(), gy.data_ptr(), gw.data_ptr(), ws.data_ptr(),
                                                                      ws.numel(), _stream(x)), (B, Cin, H, W, Cout, KH, KW, stride)),
                   "mvf_conv2d_wgrad")
        return gw
    from . import conv
    conv.require_fallback("weight gradient of conv2d(x=%s, weight=%s, stride=%s, padding=%s)" % (tuple(x.shape), tuple(wshape), stride, pad))
    conv.stats["cudnn_wgrad"] = conv.stats.get("cudnn_wgrad", 0) + 1
    w = torch.empty(wshape, device=x.device, dtype=x.dtype).contiguous(memory_format=torch.channels_last)
    _, gw, _ = torch.ops.aten.convolution_backward(gy, x, w, None, list(_pair(stride)), [pad, pad], [1, 1], False, [0, 0], 1,
                                                   [False, True, False])
    return gw


# ---- fp32-class arithmetic on the same tensor-core kernels ("3xTF32") ------------------------------------------------
# The tensor core reads the top 19 bits of an fp32 operand.  Splitting an operand into hi = round-to-tf32(v) and the exact
# remainder lo = v - hi (|lo| <= 2^-11 |v|) and summing the three products a_hi*b_hi + a_lo*b_hi + a_hi*b_lo in the fp32
# accumulator leaves a relative error of ~2^-21 per product: the arithmetic class of an fp32 convolution, which is what
# the reference runs on the CPU and with cudnn.allow_tf32 = False.  MVF_CONV_PRECISION=3xtf32 (or precision("3xtf32"))
# switches every tcgen05 convolution (forward, data gradient, weight gradient) to it; it costs 3 launches per product
# and exists for the parity tests that hold the networks to the reference at north_star's 1e-3.
_precision = os.environ.get("MVF_CONV_PRECISION", "tf32")


class precision:
    """with conv_tc.precision("3xtf32"): ...   (also usable as a plain setter: conv_tc.precision.set("tf32"))"""

    def __init__(self, mode):
        assert mode in ("tf32", "3xtf32")
        self.mode = mode

    def __enter__(self):
        global _precision
        self.old, _precision = _precision, self.mode

    def __exit__(self, *a):
        global _precision
        _precision = self.old

    @staticmethod
    def set(mode):
        global _precision
        assert mode in ("tf32", "3xtf32")
        _precision = mode


def split_tf32(v):
    """(hi, lo): hi has the low 13 mantissa bits clear (round to nearest, ties away), lo = v - hi exactly"""
    v = v.detach()
    hi = ((v.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)   # same-size dtype views keep any strides
    return hi, v - hi


class _Conv2dTC3x(torch.autograd.Function):
    """The convolution of _Conv2dTC with every product evaluated as three tensor-core products (see above)."""

    @staticmethod
    def forward(ctx, x, weight, bias, pad, stride):
        Cout, Cin, KH, KW = weight.shape
        x = _as_input(x)
        xh, xl = split_tf32(x)
        wh, wl = split_tf32(weight)
        ph, pl = pack_filters(wh), pack_filters(wl)
        y = conv_forward_raw(xh, ph, None, Cout, KH, KW, pad, stride)
        y = y + conv_forward_raw(xl, ph, None, Cout, KH, KW, pad, stride)
        y = y + conv_forward_raw(xh, pl, None, Cout, KH, KW, pad, stride)
        if bias is not None:
            y = y + bias.detach().view(1, -1, 1, 1)
        launches["fprop"] += 3
        ctx.save_for_backward(x, weight)
        ctx.pad, ctx.stride, ctx.has_bias = pad, stride, bias is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors
        Cout, Cin, KH, KW = weight.shape
        pad, stride = ctx.pad, ctx.stride
        gx = gw = gb = None
        gy = _as_input(gy)
        gh, gl = split_tf32(gy)
        if ctx.needs_input_grad[0]:
            wh, wl = split_tf32(weight)
            if _pair(stride) == (1, 1) and pad <= KH - 1 and pad <= KW - 1 and Cout % 4 == 0:
                dh, dl = pack_filters(wh, dgrad=True), pack_filters(wl, dgrad=True)
                gx = conv_forward_raw(gh, dh, None, Cin, KH, KW, KH - 1 - pad)
                gx = gx + conv_forward_raw(gl, dh, None, Cin, KH, KW, KH - 1 - pad)
                gx = gx + conv_forward_raw(gh, dl, None, Cin, KH, KW, KH - 1 - pad)
            elif _pair(stride) == (2, 2) and Cout % 4 == 0 and Cin % 4 == 0 and KH <= 8 and KW <= 8:
                gx = input_grad_s2(x, gh, wh, pad) + input_grad_s2(x, gl, wh, pad) + input_grad_s2(x, gh, wl, pad)
            else:
                gx = input_grad_library(x, gy, weight, pad, stride)
        if ctx.needs_input_grad[1]:
            xh, xl = split_tf32(x)
            gw = weight_grad(xh, gh, weight.shape, pad, stride) + weight_grad(xl, gh, weight.shape, pad, stride) + \
                weight_grad(xh, gl, weight.shape, pad, stride)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = gy.sum((0, 2, 3))
        return gx, gw, gb, None, None


# ---- disparity head: Conv3x3(C, 1) on the padded decoder feature (csrc/dispconv.cu) -- an HBM-bound direct kernel in fp32 instead
# of an N = 16 tensor-core tile with 15 zero columns (and a zero-padded 4-channel copy of grad_out for its data gradient)
dispconv_enabled = os.environ.get("MVF_DISPCONV", "1") != "0"
_disp_ws = {}


def dispconv_supported(x, weight, stride, pad):
    Cout, Cin, KH, KW = weight.shape
    return (dispconv_enabled and x.is_cuda and Cout == 1 and (KH, KW) == (3, 3) and pad == 0 and _pair(stride) == (1, 1)
            and Cin % 4 == 0 and Cin <= 64 and x.shape[1] == Cin and x.shape[2] >= 3 and x.shape[3] >= 3)


class _DispConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xp, weight, bias):
        xp = _as_input(xp)
        if not _dense_cl(xp):
            xp = xp.contiguous(memory_format=torch.channels_last)
        B, C, Hp, Wp = xp.shape
        H, W = Hp - 2, Wp - 2
        launches["fprop"] += 1
        y = torch.empty(B, 1, H, W, device=xp.device, dtype=torch.float32)
        w = weight.detach().contiguous()
        _timed("fprop", 2.0 * B * H * W * C * 9, lambda: _lib.check(_lib.lib().mvf_dispconv_fwd(
            xp.data_ptr(), w.data_ptr(), None if bias is None else bias.data_ptr(), y.data_ptr(), B, C, H, W, _stream(xp)),
            "mvf_dispconv_fwd"), (B, C, Hp, Wp, 1, 3, 3, 1))
        ctx.save_for_backward(xp, weight)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        xp, weight = ctx.saved_tensors
        B, C, Hp, Wp = xp.shape
        H, W = Hp - 2, Wp - 2
        gy = gy.contiguous().float()
        L = _lib.lib()
        gx = gw = gb = None

        def wgrad():
            launches["wgrad"] += 1
            n = L.mvf_dispconv_wgrad_workspace_floats(B * H * W, C)
            key = (gy.device.index, _stream(gy))
            ws = _disp_ws.get(key)
            if ws is None or ws.numel() < n:
                ws = _disp_ws[key] = torch.empty(n, device=gy.device, dtype=torch.float32)
            g_w = torch.empty(1, C, 3, 3, device=gy.device, dtype=torch.float32)
            g_b = torch.empty(1, device=gy.device, dtype=torch.float32) if ctx.has_bias and ctx.needs_input_grad[2] else None
            _timed("wgrad", 2.0 * B * H * W * C * 9, lambda: _lib.check(L.mvf_dispconv_wgrad(
                xp.data_ptr(), gy.data_ptr(), g_w.data_ptr(), None if g_b is None else g_b.data_ptr(), ws.data_ptr(), ws.numel(),
                B, C, H, W, _stream(gy)), "mvf_dispconv_wgrad"), (B, C, Hp, Wp, 1, 3, 3, 1))
            return g_w, g_b

        fork = None
        if ctx.needs_input_grad[1] and ctx.needs_input_grad[0] and wgrad_stream_enabled and timing is None:
            cur = torch.cuda.current_stream(gy.device)
            fork = _companion(cur)
            fork.wait_stream(cur)
            with torch.cuda.stream(fork):
                gw, gb = wgrad()
        if ctx.needs_input_grad[0]:
            launches["dgrad"] += 1
            gx = torch.empty(B, Hp, Wp, C, device=gy.device, dtype=torch.float32).permute(0, 3, 1, 2)
            w = weight.detach().contiguous()
            _timed("dgrad", 2.0 * B * H * W * C * 9, lambda: _lib.check(L.mvf_dispconv_dgrad(
                gy.data_ptr(), w.data_ptr(), gx.data_ptr(), B, C, H, W, _stream(gy)), "mvf_dispconv_dgrad"),
                (B, 1, H, W, C, 3, 3, 1))
        if fork is not None:
            cur.wait_stream(fork)
        elif ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            gw, gb = wgrad()
        return gx, gw, gb


def conv2d(x, weight, bias=None, stride=1, padding=0, act=None):
    """act: None | "relu" | "elu" -- applied in the kernel's epilogue (its backward uses the saved output)."""
    pad = padding if isinstance(padding, int) else padding[0]
    st = _pair(stride)
    st = st[0] if st[0] == st[1] else st
    if act is None and dispconv_supported(x, weight, st, pad):
        return _DispConv.apply(x, weight, bias)      # fp32 arithmetic in either precision mode
    if _precision == "3xtf32":
        y = _Conv2dTC3x.apply(x, weight, bias, pad, st)
        if act == "relu":
            return torch.relu(y)
        if act == "elu":
            return torch.nn.functional.elu(y)
        return y
    return _Conv2dTC.apply(x, weight, bias, pad, st, ACT[act])


def stem7x7s2_supported(x, weight, stride, padding):
    Cout, Cin, KH, KW = weight.shape
    return (x.is_cuda and (KH, KW) == (7, 7) and _pair(stride) == (2, 2) and _pair(padding) == (3, 3) and Cin <= 8 and
            x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0 and not x.requires_grad)


def stem7x7s2(x, weight, bias=None):
    """The 7x7 stride-2 pad-3 stem on image-like inputs (3 or 6 channels), row-packed: a channels-last pixel has only 4
    (8) floats, so a filter ROW (7 taps + 1 zero tap = 8 pixels) is ONE contiguous 32- (64-) float K block.  The padded
    image is viewed as [B, 8*C', H+6, Wo] whose column stride is 2 pixels (overlapping windows -- fine for TMA) and the
    filter as [Cout, 8*C', 7, 1]: 7 (14) pipeline stages instead of 49, no zero-channel waste, and the weight gradient
    comes out of the same tcgen05 wgrad kernel (the reshaping of the filter is plain autograd)."""
    B, Cin, H, W = x.shape
    Cout = weight.shape[0]
    Cp = 4 if Cin <= 4 else 8
    Hp, Wp, Wo = H + 6, W + 8, W // 2
    xp = torch.zeros(B, Hp, Wp, Cp, device=x.device, dtype=torch.float32)
    xp[:, 3:3 + H, 3:3 + W, :Cin] = x.permute(0, 2, 3, 1)
    fake = torch.as_strided(xp, (B, 8 * Cp, Hp, Wo), (Hp * Wp * Cp, 1, Wp * Cp, 2 * Cp))
    wrow = weight.new_zeros(Cout, 8, Cp, 7)
    wrow[:, :7, :Cin] = weight.permute(0, 3, 1, 2)          # [co, kw, c, kh]
    return conv2d(fake, wrow.reshape(Cout, 8 * Cp, 7, 1), bias, stride=(2, 1), padding=0)
