"""Drop-in for the reference's networks/IFRNet.py (inference path used by train.py:373-441 under no_grad): the frozen
video-frame-interpolation network that supplies bidirectional flows, the merge mask and the synthesized middle frame.
Same class / function names (`IFRNet`, `warp`), constructor signature and state_dict keys (`encoder.pyramid{k}.{i}.{j}.*`,
`decoder{k}.convblock.*`), so `weights/IFRNet_{L,S}_*.pth` load unchanged.  The VFI training losses of train_vfi.py are
outside the hot path and are not provided: passing `imgt` raises."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import warp_ops
from ..conv import Conv2d


def warp(img, flow):
    """backward warp by a pixel-unit flow, border padding, align_corners=True (IFRNet.py:7-15): mvf_flow_warp_fwd on CUDA"""
    return warp_ops.flow_warp(img, flow)


def resize(x, scale_factor, mul=(1.0, 1.0)):
    """IFRNet.py:117-118 (+ the flow multiplier of its call sites in the same kernel)"""
    return warp_ops.resize_bilinear(x, scale_factor=scale_factor, align_corners=False, mul=mul)


class PReLU(nn.PReLU):
    """nn.PReLU with the same parameter (`weight`); on CUDA under no_grad the channels-last kernel mvf_prelu_cl_fwd"""

    def forward(self, x, res=None):
        return warp_ops.prelu(x, self.weight, res)


def _fusable(x):
    """inference on the single-pass TF32 tensor-core path (the fp32-class 3xTF32 mode keeps the unfused modules)"""
    from .. import conv, conv_tc
    return x.is_cuda and conv.get_backend() == "tcgen05" and conv_tc._precision == "tf32" and not torch.is_grad_enabled()


class ConvPReLU(nn.Sequential):
    """Conv2d -> PReLU with the reference's state_dict keys (`0.weight`, `0.bias`, `1.weight`); on CUDA under no_grad the PReLU
    runs in the convolution's epilogue (mvf_conv2d_forward_prelu)"""

    def forward(self, x):
        from .. import conv, conv_tc
        c, a = self[0], self[1]
        if _fusable(x) and x.shape[1] % 4 == 0 and conv_tc.supported(x, c.weight, c.stride, c.padding, c.dilation, c.groups):
            conv.stats["tcgen05"] += 1
            return conv_tc.conv2d_prelu_inference(x, c.weight, c.bias, a.weight, c.stride, c.padding)
        return a(c(x))


class ConvTranspose2d(nn.ConvTranspose2d):
    """nn.ConvTranspose2d(k=4, stride 2, padding 1) of the decoders (IFRNet.py:194); on CUDA under no_grad: the tcgen05 stride-2
    data-gradient kernel with the bias in its epilogue (mvf_conv_transpose2d_s2_fwd)"""

    def forward(self, x):
        from .. import conv, conv_tc
        if (_fusable(x) and self.stride == (2, 2) and self.padding[0] == self.padding[1] and self.output_padding == (0, 0) and
                self.groups == 1 and self.dilation == (1, 1) and conv_tc.dgrad_s2_enabled):
            conv.stats["tcgen05"] += 1
            return conv_tc.conv_transpose2d_s2_inference(x, self.weight, self.bias, self.padding[0])
        conv.stats["cudnn"] += 1
        return super().forward(x)


def convrelu(cin, cout, k=3, stride=1, padding=1):
    return ConvPReLU(Conv2d(cin, cout, k, stride, padding, bias=True), PReLU(cout))


class ResBlock(nn.Module):
    """five 3x3 convs; the 2nd and 4th only refresh the last `side` channels (IFRNet.py:122-150)"""

    def __init__(self, channels, side):
        super().__init__()
        self.side_channels = side
        self.conv1 = convrelu(channels, channels)
        self.conv2 = convrelu(side, side)
        self.conv3 = convrelu(channels, channels)
        self.conv4 = convrelu(side, side)
        self.conv5 = Conv2d(channels, channels, 3, 1, 1, bias=True)
        self.prelu = PReLU(channels)

    def _side(self, conv, y):
        s = self.side_channels
        return torch.cat([y[:, :-s], conv(y[:, -s:])], 1)

    def forward(self, x):
        y = self._side(self.conv2, self.conv1(x))
        y = self._side(self.conv4, self.conv3(y))
        return self.prelu(self.conv5(y), x)


class _Encoder(nn.Module):
    def __init__(self, widths, first_kernel):
        super().__init__()
        cin = 3
        for k, c in enumerate(widths, 1):
            first = convrelu(cin, c, first_kernel, 2, first_kernel // 2) if k == 1 else convrelu(cin, c, 3, 2, 1)
            setattr(self, "pyramid%d" % k, nn.Sequential(first, convrelu(c, c, 3, 1, 1)))
            cin = c

    def forward(self, img):
        f1 = self.pyramid1(img)
        f2 = self.pyramid2(f1)
        f3 = self.pyramid3(f2)
        return f1, f2, f3, self.pyramid4(f3)


class _Decoder(nn.Module):
    """convrelu -> ResBlock -> transposed conv x2; the coarsest one takes the time embedding, the others the warped
    pyramid features and the current flows (IFRNet.py:153-211)"""

    def __init__(self, cin, mid, side, cout):
        super().__init__()
        self.convblock = nn.Sequential(convrelu(cin, mid), ResBlock(mid, side), ConvTranspose2d(mid, cout, 4, 2, 1, bias=True))

    def forward(self, *inputs):
        return self.convblock(torch.cat(inputs, 1))


# pyramid widths, ResBlock side channels and first kernel of the two published sizes (IFRNet.py:153-330)
_SIZES = {"large": ([64, 96, 144, 192], 64, 7), "small": ([24, 36, 54, 72], 24, 3)}


class IFRNet(nn.Module):
    def __init__(self, scale="large"):
        super().__init__()
        w, side, k1 = _SIZES[scale]
        self.encoder = _Encoder(w, k1)
        # decoder k consumes level-k features of both frames (+ the feature predicted for the middle frame) and emits
        # 4 flow channels + the next finer middle-frame feature (decoder1: 4 flows + mask + 3 residual channels)
        self.decoder4 = _Decoder(2 * w[3] + 1, 2 * w[3], side, w[2] + 4)
        self.decoder3 = _Decoder(3 * w[2] + 4, 3 * w[2], side, w[1] + 4)
        self.decoder2 = _Decoder(3 * w[1] + 4, 3 * w[1], side, w[0] + 4)
        self.decoder1 = _Decoder(3 * w[0] + 4, 3 * w[0], side, 8)

    def forward(self, img0, img1, embt, imgt=None, scale_factor=(1.0, 0.5), onlyFlow=False):
        if imgt is not None:
            raise NotImplementedError("the VFI training losses (train_vfi.py) are outside the depth-training hot path")
        _, _, H, W = img0.shape
        if H == 320 and W == 1024:
            scale_factor = (0.6, 0.3125)
        mean_ = torch.cat([img0, img1], 2).mean(1, keepdim=True).mean(2, keepdim=True).mean(3, keepdim=True)
        img0, img1 = img0 - mean_, img1 - mean_
        size = (int(H * scale_factor[0]), int(W * scale_factor[1]))
        f0 = self.encoder(warp_ops.resize_bilinear(img0, size=size))
        f1 = self.encoder(warp_ops.resize_bilinear(img1, size=size))
        b, c, h, w = f0[3].shape
        out = self.decoder4(f0[3], f1[3], embt.repeat(1, 1, h, w))
        flow0, flow1, ft = out[:, 0:2], out[:, 2:4], out[:, 4:]
        for level, dec in ((2, self.decoder3), (1, self.decoder2), (0, self.decoder1)):
            out = dec(ft, warp(f0[level], flow0), warp(f1[level], flow1), flow0, flow1)
            flow0 = out[:, 0:2] + resize(flow0, 2.0, mul=(2.0, 2.0))
            flow1 = out[:, 2:4] + resize(flow1, 2.0, mul=(2.0, 2.0))
            ft = out[:, 4:]
        mask = torch.sigmoid(out[:, 4:5])
        rescale = (1.0 / scale_factor[1], 1.0 / scale_factor[0])   # (x, y) displacement per full-resolution pixel
        flow0 = warp_ops.resize_bilinear(flow0, size=(H, W), mul=rescale)
        flow1 = warp_ops.resize_bilinear(flow1, size=(H, W), mul=rescale)
        mask = warp_ops.resize_bilinear(mask, size=(H, W))
        if onlyFlow:
            return flow0, flow1, mask
        merged = mask * warp(img0, flow0) + (1 - mask) * warp(img1, flow1)
        return torch.clamp(merged + mean_, 0, 1), flow0, flow1, mask
