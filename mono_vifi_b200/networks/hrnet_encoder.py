"""HRNet-W18 feature extractor with the reference's parameter names (hrnet_encoder.py:294-498 + hrnet_config.py HRNET_18),
so `weights/HRNet_W18_C_*.pth.tar` and reference checkpoints load unchanged.  Convolutions go through conv.Conv2d
(tcgen05 kernels on CUDA).  Written from the architecture, not from the reference's code: stages are described by a
small table and built by three helpers (residual unit, branch, exchange unit)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import warp_ops
from ..bn_act import bn_act
from ..conv import Conv2d

# (modules, blocks per branch, channels per branch, block type) for stages 1..4 of HRNet-W18  (hrnet_config.py:119-153)
HRNET18 = {
    1: dict(modules=1, blocks=[4], channels=[64], kind="bottleneck"),
    2: dict(modules=1, blocks=[4, 4], channels=[18, 36], kind="basic"),
    3: dict(modules=4, blocks=[4, 4, 4], channels=[18, 36, 72], kind="basic"),
    4: dict(modules=3, blocks=[4, 4, 4, 4], channels=[18, 36, 72, 144], kind="basic"),
}


def _conv_bn(cin, cout, k, stride, relu):
    layers = [Conv2d(cin, cout, k, stride, k // 2, bias=False), nn.BatchNorm2d(cout)]
    if relu:
        layers.append(nn.ReLU())
    return nn.Sequential(*layers)


class BasicUnit(nn.Module):
    """two 3x3 convs + identity (hrnet_encoder.py:58-94); names conv1/bn1/conv2/bn2/downsample"""
    expansion = 1

    def __init__(self, cin, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = Conv2d(cin, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU()
        self.conv2 = Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample

    def forward(self, x):
        identity = x if self.downsample is None else self.downsample(x)
        return bn_act(self.bn2, self.conv2(bn_act(self.bn1, self.conv1(x))), identity)


class BottleneckUnit(nn.Module):
    """1x1 - 3x3 - 1x1 (x4) + identity (hrnet_encoder.py:97-139)"""
    expansion = 4

    def __init__(self, cin, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = Conv2d(cin, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = Conv2d(planes, planes, 3, stride, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU()
        self.downsample = downsample

    def forward(self, x):
        identity = x if self.downsample is None else self.downsample(x)
        y = bn_act(self.bn2, self.conv2(bn_act(self.bn1, self.conv1(x))))
        return bn_act(self.bn3, self.conv3(y), identity)


_UNITS = {"basic": BasicUnit, "bottleneck": BottleneckUnit}


def _chain(unit, cin, planes, n):
    """n residual units; the first one projects the identity when the width changes"""
    ds = None
    if cin != planes * unit.expansion:
        ds = nn.Sequential(Conv2d(cin, planes * unit.expansion, 1, 1, bias=False), nn.BatchNorm2d(planes * unit.expansion))
    mods = [unit(cin, planes, 1, ds)]
    mods += [unit(planes * unit.expansion, planes) for _ in range(n - 1)]
    return nn.Sequential(*mods)


class HighResolutionModule(nn.Module):
    """parallel branches followed by the all-to-all exchange (hrnet_encoder.py:142-287): branch j -> resolution i goes
    through strided 3x3 convs (j < i), identity (j == i) or a 1x1 conv + bilinear upsampling (j > i)."""

    def __init__(self, channels, blocks, unit):
        super().__init__()
        n = len(channels)
        self.num_branches = n
        self.branches = nn.ModuleList([_chain(unit, c, c // unit.expansion, b) for c, b in zip(channels, blocks)])
        self.fuse_layers = None
        if n > 1:
            rows = []
            for i in range(n):
                row = []
                for j in range(n):
                    if j > i:
                        row.append(_conv_bn(channels[j], channels[i], 1, 1, relu=False))
                    elif j == i:
                        row.append(None)
                    else:
                        steps = [_conv_bn(channels[j], channels[j], 3, 2, relu=True) for _ in range(i - j - 1)]
                        steps.append(_conv_bn(channels[j], channels[i], 3, 2, relu=False))
                        row.append(nn.Sequential(*steps))
                rows.append(nn.ModuleList(row))
            self.fuse_layers = nn.ModuleList(rows)
        self.relu = nn.ReLU()

    def forward(self, xs):
        xs = [br(x) for br, x in zip(self.branches, xs)]
        if self.num_branches == 1:
            return xs
        outs = []
        for i, row in enumerate(self.fuse_layers):
            y = xs[0] if i == 0 else row[0](xs[0])
            for j in range(1, self.num_branches):
                if j == i:
                    y = y + xs[j]
                elif j > i:
                    y = y + warp_ops.resize_bilinear(row[j](xs[j]), size=xs[i].shape[-2:], align_corners=True)
                else:
                    y = y + row[j](xs[j])
            outs.append(self.relu(y))
        return outs


class HighResolutionNet(nn.Module):
    def __init__(self, table=HRNET18):
        super().__init__()
        self.conv1 = Conv2d(3, 64, 3, 2, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.conv2 = Conv2d(64, 64, 3, 2, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU()
        s1 = table[1]
        unit1 = _UNITS[s1["kind"]]
        self.layer1 = _chain(unit1, 64, s1["channels"][0], s1["blocks"][0])
        prev = [s1["channels"][0] * unit1.expansion]
        for s in (2, 3, 4):
            cfg = table[s]
            unit = _UNITS[cfg["kind"]]
            chans = [c * unit.expansion for c in cfg["channels"]]
            setattr(self, "transition%d" % (s - 1), self._transition(prev, chans))
            setattr(self, "stage%d" % s, nn.Sequential(*[HighResolutionModule(chans, cfg["blocks"], unit)
                                                          for _ in range(cfg["modules"])]))
            prev = chans
        self.num_branches = [len(table[s]["channels"]) for s in (2, 3, 4)]

    @staticmethod
    def _transition(prev, cur):
        """new branches are created from the lowest-resolution one by strided 3x3 convs; widths are adapted by a 3x3 conv
        (hrnet_encoder.py:355-389)"""
        layers = []
        for i, c in enumerate(cur):
            if i < len(prev):
                layers.append(_conv_bn(prev[i], c, 3, 1, relu=True) if c != prev[i] else None)
            else:
                steps, n = [], i + 1 - len(prev)
                for j in range(n):
                    steps.append(_conv_bn(prev[-1], c if j == n - 1 else prev[-1], 3, 2, relu=True))
                layers.append(nn.Sequential(*steps))
        return nn.ModuleList(layers)

    def forward(self, x):
        x = self.relu(self.bn1(self.conv1(x)))
        outputs = [x]
        x = self.layer1(self.relu(self.bn2(self.conv2(x))))
        ys = [x]
        for s in (2, 3, 4):
            tr = getattr(self, "transition%d" % (s - 1))
            xs = []
            for i, t in enumerate(tr):
                src = ys[i] if i < len(ys) else ys[-1]
                xs.append(src if t is None else t(src))
            ys = getattr(self, "stage%d" % s)(xs)
        return outputs + ys  # 64 @ 1/2, then 18 / 36 / 72 / 144 channels @ 1/4 .. 1/32


def hrnet18(pretrained=True, progress=True):
    model = HighResolutionNet(HRNET18)
    if pretrained:  # same file the reference expects (hrnet_encoder.py:505-507)
        loaded = torch.load("./weights/HRNet_W18_C_cosinelr_cutmix_300epoch.pth.tar", map_location="cpu")
        own = model.state_dict()
        model.load_state_dict({k: v for k, v in loaded.items() if k in own})
    model.num_ch_enc = [64, 18, 36, 72, 144]
    return model
