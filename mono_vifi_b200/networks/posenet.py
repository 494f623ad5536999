"""Drop-in for the reference's networks/posenet.py: ResNet pose encoder (6-channel stem) + PoseDecoder."""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

from ..conv import Conv2d
from ..bn_act import bn_act
from .resnet import ResNet, load_imagenet


def resnet_multiimage_input(num_layers, pretrained=False, num_input_images=1):
    """posenet.py:35-52"""
    assert num_layers in [18, 50], "Can only run with 18 or 50 layer resnet"
    model = ResNet(num_layers, in_channels=num_input_images * 3)
    if pretrained:
        load_imagenet(model, num_layers, num_input_images)
    return model


class ResnetEncoder(nn.Module):
    """posenet.py:55-97"""

    def __init__(self, num_layers, pretrained, num_input_images=1):
        super().__init__()
        self.num_ch_enc = np.array([64, 64, 128, 256, 512])
        if num_layers not in (18, 34, 50, 101, 152):
            raise ValueError("{} is not a valid number of resnet layers".format(num_layers))
        if num_input_images > 1:
            self.encoder = resnet_multiimage_input(num_layers, pretrained, num_input_images)
        else:
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


class PoseDecoder(nn.Module):
    """posenet.py:100-137"""

    def __init__(self, num_ch_enc, num_input_features, num_frames_to_predict_for=None, stride=1):
        super().__init__()
        self.num_ch_enc = num_ch_enc
        self.num_input_features = num_input_features
        if num_frames_to_predict_for is None:
            num_frames_to_predict_for = num_input_features - 1
        self.num_frames_to_predict_for = num_frames_to_predict_for
        self.convs = OrderedDict()
        self.convs[("squeeze")] = Conv2d(int(self.num_ch_enc[-1]), 256, 1)
        self.convs[("pose", 0)] = Conv2d(num_input_features * 256, 256, 3, stride, 1)
        self.convs[("pose", 1)] = Conv2d(256, 256, 3, stride, 1)
        self.convs[("pose", 2)] = Conv2d(256, 6 * num_frames_to_predict_for, 1)
        self.relu = nn.ReLU()
        self.net = nn.ModuleList(list(self.convs.values()))

    def forward(self, input_features):
        last_features = [f[-1] for f in input_features]
        cat_features = [self.relu(self.convs["squeeze"](f)) for f in last_features]
        cat_features = torch.cat(cat_features, 1)
        out = cat_features
        for i in range(3):
            out = self.convs[("pose", i)](out)
            if i != 2:
                out = self.relu(out)
        out = out.mean(3).mean(2)
        out = 0.01 * out.view(-1, self.num_frames_to_predict_for, 1, 6)
        return out[..., :3], out[..., 3:]
