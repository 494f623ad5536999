"""Drop-in for the reference's networks/IFRNet.py (inference path used by train.py:373-441 under no_grad): the frozen
video-frame-interpolation network that supplies bidirectional flows, the merge mask and the synthesized middle frame.
Same class / function names (`IFRNet`, `warp`), constructor signature and state_dict keys (`encoder.pyramid{k}.{i}.{j}.*`,
`decoder{k}.convblock.*`), so `weights/IFRNet_{L,S}_*.pth` load unchanged.  The VFI training losses of train_vfi.py are
outside the hot path and are not provided: passing `imgt` raises."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..conv import Conv2d


def warp(img, flow):
    """backward warp by a pixel-unit flow, border padding, align_corners=True (IFRNet.py:7-15)"""
    B, _, H, W = flow.shape
    xx = torch.linspace(-1.0, 1.0, W, device=flow.device, dtype=flow.dtype).view(1, 1, 1, W).expand(B, -1, H, -1)
    yy = torch.linspace(-1.0, 1.0, H, device=flow.device, dtype=flow.dtype).view(1, 1, H, 1).expand(B, -1, -1, W)
    grid = torch.cat([xx + flow[:, 0:1] / ((W - 1.0) / 2.0), yy + flow[:, 1:2] / ((H - 1.0) / 2.0)], 1).to(img)
    return F.grid_sample(img, grid.permute(0, 2, 3, 1), mode="bilinear", padding_mode="border", align_corners=True)


def resize(x, scale_factor):
    return F.interpolate(x, scale_factor=scale_factor, mode="bilinear", align_corners=False)


def convrelu(cin, cout, k=3, stride=1, padding=1):
    return nn.Sequential(Conv2d(cin, cout, k, stride, padding, bias=True), nn.PReLU(cout))


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
        self.prelu = nn.PReLU(channels)

    def _side(self, conv, y):
        s = self.side_channels
        return torch.cat([y[:, :-s], conv(y[:, -s:])], 1)

    def forward(self, x):
        y = self._side(self.conv2, self.conv1(x))
        y = self._side(self.conv4, self.conv3(y))
        return self.prelu(x + self.conv5(y))


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
        self.convblock = nn.Sequential(convrelu(cin, mid), ResBlock(mid, side), nn.ConvTranspose2d(mid, cout, 4, 2, 1, bias=True))

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
        f0 = self.encoder(F.interpolate(img0, size=size, mode="bilinear", align_corners=False))
        f1 = self.encoder(F.interpolate(img1, size=size, mode="bilinear", align_corners=False))
        b, c, h, w = f0[3].shape
        out = self.decoder4(f0[3], f1[3], embt.repeat(1, 1, h, w))
        flow0, flow1, ft = out[:, 0:2], out[:, 2:4], out[:, 4:]
        for level, dec in ((2, self.decoder3), (1, self.decoder2), (0, self.decoder1)):
            out = dec(ft, warp(f0[level], flow0), warp(f1[level], flow1), flow0, flow1)
            flow0 = out[:, 0:2] + 2.0 * resize(flow0, 2.0)
            flow1 = out[:, 2:4] + 2.0 * resize(flow1, 2.0)
            ft = out[:, 4:]
        mask = torch.sigmoid(out[:, 4:5])
        rescale = flow0.new_tensor([1.0 / scale_factor[1], 1.0 / scale_factor[0]]).view(1, 2, 1, 1)
        flow0 = F.interpolate(flow0, size=(H, W), mode="bilinear", align_corners=False) * rescale
        flow1 = F.interpolate(flow1, size=(H, W), mode="bilinear", align_corners=False) * rescale
        mask = F.interpolate(mask, size=(H, W), mode="bilinear", align_corners=False)
        if onlyFlow:
            return flow0, flow1, mask
        merged = mask * warp(img0, flow0) + (1 - mask) * warp(img1, flow1)
        return torch.clamp(merged + mean_, 0, 1), flow0, flow1, mask
