"""Drop-in for the reference's networks/monodepth2.py: ResNet depth encoder + U-Net disparity decoder.
Same class names, constructor signatures, attributes (num_ch_enc, features, outputs) and state_dict keys
(`encoder.conv1.weight` ..., `decoder.{i}.conv.conv.weight`)."""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

from ..layers import ConvBlock, Conv3x3, upsample
from ..bn_act import bn_act
from .resnet import ResNet, load_imagenet


class DepthEncoder(nn.Module):
    """monodepth2.py:11-45"""

    def __init__(self, num_layers, pretrained):
        super().__init__()
        self.num_ch_enc = np.array([64, 64, 128, 256, 512])
        self.encoder = ResNet(num_layers)
        if pretrained:
            load_imagenet(self.encoder, num_layers)
        if num_layers > 34:
            self.num_ch_enc[1:] *= 4

    def forward(self, input_image):
        self.features = []
        x = (input_image - 0.45) / 0.225
        x = self.encoder.conv1(x)
        self.features.append(bn_act(self.encoder.bn1, x))  # bn1 + relu
        self.features.append(self.encoder.layer1(self.encoder.maxpool(self.features[-1])))
        self.features.append(self.encoder.layer2(self.features[-1]))
        self.features.append(self.encoder.layer3(self.features[-1]))
        self.features.append(self.encoder.layer4(self.features[-1]))
        return self.features


class DepthDecoder(nn.Module):
    """monodepth2.py:48-96"""

    def __init__(self, num_ch_enc, scales=range(4), num_output_channels=1, use_skips=True):
        super().__init__()
        self.num_output_channels = num_output_channels
        self.use_skips = use_skips
        self.scales = scales
        self.num_ch_enc = num_ch_enc
        self.num_ch_dec = np.array([16, 32, 64, 128, 256])
        self.convs = OrderedDict()
        for i in range(4, -1, -1):
            num_ch_in = self.num_ch_enc[-1] if i == 4 else self.num_ch_dec[i + 1]
            self.convs[("upconv", i, 0)] = ConvBlock(num_ch_in, self.num_ch_dec[i])
            num_ch_in = self.num_ch_dec[i]
            if self.use_skips and i > 0:
                num_ch_in += self.num_ch_enc[i - 1]
            self.convs[("upconv", i, 1)] = ConvBlock(num_ch_in, self.num_ch_dec[i])
        for s in self.scales:
            self.convs[("dispconv", s)] = Conv3x3(self.num_ch_dec[s], self.num_output_channels)
        self.decoder = nn.ModuleList(list(self.convs.values()))
        self.sigmoid = nn.Sigmoid()

    def forward(self, input_features):
        self.outputs = {}
        x = input_features[-1]
        for i in range(4, -1, -1):
            x = self.convs[("upconv", i, 0)](x)
            skip = input_features[i - 1] if (self.use_skips and i > 0) else None
            # upsample + concat + (the ConvBlock's) reflection pad in one pass, then conv + ELU  (monodepth2.py:86-93)
            x = self.convs[("upconv", i, 1)].forward_upcat(x, skip, upsample=True)
            if i in self.scales:
                self.outputs[("disp", i)] = self.sigmoid(self.convs[("dispconv", i)](x))
        return self.outputs
