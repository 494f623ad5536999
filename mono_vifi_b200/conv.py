"""Convolution entry point of the network stacks.

`Conv2d` keeps nn.Conv2d's parameters / state_dict keys (`weight`, `bias`), so checkpoints of the reference load
unchanged, and routes the arithmetic through `conv2d()`:

  backend "tcgen05"  hand-written sm_100a implicit-GEMM kernels (mono_vifi_b200/csrc/conv_tc.cu, conv_wgrad.cu) on
                     channels-last activations -- the default on a CUDA device;
  backend "cudnn"    torch's F.conv2d (cuDNN, a LIBRARY baseline -- what the reference itself runs on), kept for
                     A/B timing (MVF_CONV_BACKEND=cudnn) and for CPU tensors (the host-side tests).

Every call is counted in `stats` so the bench can report how much of the conv work ran on which path.  A CUDA tensor
whose shape the tcgen05 kernels do not cover (grouped or dilated convolutions, asymmetric padding, strides above 2) RAISES
`UnsupportedConvolution`: the product path has no library fallback.  MVF_LIBRARY_FALLBACK=1 turns the error into a counted
cuDNN call (porting aid for other architectures; none of the five BASELINE configurations needs it).
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

_backend = os.environ.get("MVF_CONV_BACKEND", "tcgen05")
library_fallback = os.environ.get("MVF_LIBRARY_FALLBACK", "0") == "1"
stats = {"tcgen05": 0, "cudnn": 0}


class UnsupportedConvolution(RuntimeError):
    pass


def require_fallback(what):
    """Called where a shape would leave the tcgen05 kernels: raises unless the counted library fallback was asked for."""
    if not library_fallback:
        raise UnsupportedConvolution("%s is not covered by the tcgen05 convolution kernels and the product path has no library "
                                     "fallback (MVF_LIBRARY_FALLBACK=1 allows a counted cuDNN call; MVF_CONV_BACKEND=cudnn runs "
                                     "the whole network on the library for A/B timing)" % what)


def set_backend(name):
    global _backend
    if name not in ("cudnn", "tcgen05"):
        raise ValueError("unknown conv backend %r" % (name,))
    _backend = name


def get_backend():
    return _backend


def _activate(y, act):
    if act == "relu":
        return F.relu(y)
    if act == "elu":
        return F.elu(y)
    return y


def conv2d(x, weight, bias=None, stride=1, padding=0, dilation=1, groups=1, act=None):
    if _backend == "tcgen05" and x.is_cuda:
        from . import conv_tc
        if groups == 1 and act is None and conv_tc.stem7x7s2_supported(x, weight, stride, padding):
            stats["tcgen05"] += 1
            return conv_tc.stem7x7s2(x, weight, bias)
        cin = weight.shape[1]
        if cin % 4 and groups == 1:
            # image-like inputs (3 or 6 channels): zero channels up to a 16-byte pixel; autograd slices the gradient back
            extra = 4 - cin % 4
            x = F.pad(x, (0, 0, 0, 0, 0, extra))
            weight = F.pad(weight, (0, 0, 0, 0, 0, extra))
        if conv_tc.supported(x, weight, stride, padding, dilation, groups):
            stats["tcgen05"] += 1
            cout = weight.shape[0]
            if cout % 4 and cout > 4:
                # HRNet-W18's 18-channel branch: zero output channels up to a 16-byte pixel, so that the backward's data /
                # weight gradients stay on the tensor-core kernels too; the result is the leading-channel view
                extra = 4 - cout % 4
                weight = F.pad(weight, (0, 0, 0, 0, 0, 0, 0, extra))
                bias = None if bias is None else F.pad(bias, (0, extra))
                return conv_tc.conv2d(x, weight, bias, stride, padding, act=act)[:, :cout]
            return conv_tc.conv2d(x, weight, bias, stride, padding, act=act)
        require_fallback("conv2d(x=%s, weight=%s, stride=%s, padding=%s, dilation=%s, groups=%s)" %
                         (tuple(x.shape), tuple(weight.shape), stride, padding, dilation, groups))
    stats["cudnn"] += 1
    return _activate(F.conv2d(x, weight, bias, stride, padding, dilation, groups), act)


class Conv2d(nn.Conv2d):
    def forward(self, x):
        if self.padding_mode != "zeros":
            return super().forward(x)
        return conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
