"""Drop-in for the reference's networks/DHRNet.py: HRNet-W18 depth encoder + dense multi-scale-fusion decoder
(DHRNet.py:9-145).  Same class names, constructor signatures, attributes and state_dict keys (`encoder.*`,
`decoder.{i}.conv.conv.*`)."""
from collections import OrderedDict

import numpy as np
import torch.nn as nn

from ..layers import ConvBlock, ConvBlock1x1, Conv3x3, upsample
from .hrnet_encoder import hrnet18


class DepthEncoder(nn.Module):
    def __init__(self, num_layers, pretrained):
        super().__init__()
        assert num_layers == 18
        self.encoder = hrnet18(pretrained)
        self.num_ch_enc = np.array(self.encoder.num_ch_enc)

    def forward(self, x):
        self.features = self.encoder((x - 0.45) / 0.225)
        return self.features


class DepthDecoder(nn.Module):
    """Levels 1..4 (1/4 .. 1/32) are refined in three rounds; in each round every level is convolved (ConvBlock), the
    coarser ones are upsampled, squeezed by a 1x1 ConvBlock and summed into the finer ones; the coarsest level is dropped
    after each round.  Then level 1 joins the 1/2-resolution stem feature and two ConvBlocks lead to the disparity."""

    def __init__(self, num_ch_enc, scales=range(4), num_output_channels=1):
        super().__init__()
        self.num_output_channels = num_output_channels
        self.scales = scales
        self.num_ch_enc = c = num_ch_enc
        self.convs = OrderedDict()
        # registration order = the reference's, so that `decoder.{i}` indices match (DHRNet.py:37-66)
        for rnd, top in ((0, 4), (1, 3), (2, 2)):
            for lv in range(1, top + 1):
                self.convs[("parallel_conv"), rnd, lv] = ConvBlock(c[lv], c[lv])
            for src in range(2, top + 1):
                for dst in range(src - 1, 0, -1):
                    self.convs[("conv1x1", rnd, src * 10 + dst)] = ConvBlock1x1(c[src], c[dst])
        self.convs[("parallel_conv"), 3, 0] = ConvBlock(c[0], c[0])
        self.convs[("parallel_conv"), 3, 1] = ConvBlock(c[1], c[1])
        self.convs[("conv1x1", 3, 10)] = ConvBlock1x1(c[1], c[0])
        self.convs[("parallel_conv"), 4, 0] = ConvBlock(c[0], 32)
        self.convs[("parallel_conv"), 5, 0] = ConvBlock(32, 16)
        self.convs[("dispconv", 0)] = Conv3x3(16, self.num_output_channels)
        self.decoder = nn.ModuleList(list(self.convs.values()))
        self.sigmoid = nn.Sigmoid()

    def forward(self, input_features):
        self.outputs = {}
        level = {lv: input_features[lv] for lv in range(1, 5)}
        for rnd, top in ((0, 4), (1, 3), (2, 2)):
            d = {lv: self.convs[("parallel_conv"), rnd, lv](level[lv]) for lv in range(1, top + 1)}
            fused = {}
            for dst in range(1, top):
                y = d[dst]
                for src in range(dst + 1, top + 1):
                    y = y + self.convs[("conv1x1", rnd, src * 10 + dst)](upsample(d[src], 2 ** (src - dst)))
                fused[dst] = y
            level = fused
        d0 = self.convs[("parallel_conv"), 3, 0](input_features[0])
        d1 = self.convs[("parallel_conv"), 3, 1](level[1])
        x = d0 + self.convs[("conv1x1", 3, 10)](upsample(d1, 2))
        x = upsample(self.convs[("parallel_conv"), 4, 0](x), 2)
        x = self.convs[("parallel_conv"), 5, 0](x)
        self.outputs[("disp", 0)] = self.sigmoid(self.convs[("dispconv", 0)](x))
        return self.outputs
