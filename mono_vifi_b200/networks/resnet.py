"""ResNet-18/34/50 feature extractor with torchvision's parameter names (conv1, bn1, layer{1..4}.{i}.conv{j},
...downsample.{0,1}, fc) and initialisation, so the reference's checkpoints (`encoder.*` keys) load unchanged.
Used by networks/monodepth2.py (DepthEncoder) and networks/posenet.py (ResnetEncoder).
Reference: monodepth2.py:16-31, posenet.py:10-52 (both wrap torchvision.models.resnet)."""
import torch.nn as nn

from ..bn_act import bn_act
from ..conv import Conv2d
from ..decoder_ops import MaxPool3s2


def _shortcut(downsample, x):
    """identity branch: nothing, or the 1x1 strided conv + bn pair (bn through the fused kernels, no activation)"""
    if downsample is None:
        return x
    if isinstance(downsample, nn.Sequential) and len(downsample) == 2 and isinstance(downsample[1], nn.modules.batchnorm._BatchNorm):
        return bn_act(downsample[1], downsample[0](x), None, relu=False)
    return downsample(x)


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        identity = _shortcut(self.downsample, x)
        out = bn_act(self.bn1, self.conv1(x))                        # bn + relu in one pass
        return bn_act(self.bn2, self.conv2(out), identity)           # bn + residual add + relu in one pass


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = Conv2d(planes, planes, 3, stride, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        identity = _shortcut(self.downsample, x)
        out = bn_act(self.bn1, self.conv1(x))
        out = bn_act(self.bn2, self.conv2(out))
        return bn_act(self.bn3, self.conv3(out), identity)


_CFG = {18: (BasicBlock, [2, 2, 2, 2]), 34: (BasicBlock, [3, 4, 6, 3]), 50: (Bottleneck, [3, 4, 6, 3]),
        101: (Bottleneck, [3, 4, 23, 3]), 152: (Bottleneck, [3, 8, 36, 3])}


class ResNet(nn.Module):
    def __init__(self, num_layers, in_channels=3, num_classes=1000):
        super().__init__()
        if num_layers not in _CFG:
            raise ValueError("{} is not a valid number of resnet layers".format(num_layers))
        block, layers = _CFG[num_layers]
        self.inplanes = 64
        self.conv1 = Conv2d(in_channels, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = MaxPool3s2()  # MaxPool2d(3, 2, 1): own channels-last kernels on CUDA, F.max_pool2d elsewhere
        self.layer1 = self._make_layer(block, 64, layers[0])
        self.layer2 = self._make_layer(block, 128, layers[1], stride=2)
        self.layer3 = self._make_layer(block, 256, layers[2], stride=2)
        self.layer4 = self._make_layer(block, 512, layers[3], stride=2)
        # never used by the depth / pose nets, kept so state_dicts match the reference's
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(512 * block.expansion, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def _make_layer(self, block, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(Conv2d(self.inplanes, planes * block.expansion, 1, stride, bias=False),
                                       nn.BatchNorm2d(planes * block.expansion))
        layers = [block(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes * block.expansion
        for _ in range(1, blocks):
            layers.append(block(self.inplanes, planes))
        return nn.Sequential(*layers)


def load_imagenet(model, num_layers, num_input_images=1):
    """`pretrained=True` of the reference downloads torchvision ImageNet weights; same here (needs network)."""
    import torch
    import torchvision.models as tvm
    names = {18: "ResNet18_Weights", 34: "ResNet34_Weights", 50: "ResNet50_Weights", 101: "ResNet101_Weights",
             152: "ResNet152_Weights"}
    loaded = getattr(tvm, names[num_layers]).IMAGENET1K_V1.get_state_dict(progress=False)
    if num_input_images > 1:  # posenet.py:48-50
        loaded['conv1.weight'] = torch.cat([loaded['conv1.weight']] * num_input_images, 1) / num_input_images
    model.load_state_dict(loaded)
