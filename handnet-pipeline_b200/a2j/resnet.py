"""Parameter containers of the A2J backbone (reference: a2j/resnet.py).

ResNet-50 with the stride on the 3x3 convolution of each Bottleneck (resnet.py:68) and a dilated, stride-1
layer4 (dilation 2 on blocks 1-2 only, resnet.py:112,142,145).  The modules are never called: the state-dict
keys (incl. the unused ``fc``) match the reference so its checkpoints load, and hn_b200.runtime reads the
tensors to drive the tcgen05 convolution kernels.  No ImageNet download is attempted (no network; the
reference's ``pretrained=True`` only matters for training).
"""
from __future__ import annotations

import torch.nn as nn

__all__ = ["ResNet", "Bottleneck", "resnet50"]


def conv3x3(cin, cout, stride=1, dilation=1):
    return nn.Conv2d(cin, cout, kernel_size=3, stride=stride, dilation=dilation, padding=dilation, bias=False)


def conv1x1(cin, cout, stride=1):
    return nn.Conv2d(cin, cout, kernel_size=1, stride=stride, bias=False)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, dilation=1):
        super().__init__()
        self.conv1, self.bn1 = conv1x1(inplanes, planes), nn.BatchNorm2d(planes)
        self.conv2, self.bn2 = conv3x3(planes, planes, stride, dilation), nn.BatchNorm2d(planes)
        self.conv3, self.bn3 = conv1x1(planes, planes * 4), nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride


class ResNet(nn.Module):
    def __init__(self, block=Bottleneck, layers=(3, 4, 6, 3), num_classes=1000):
        super().__init__()
        self.inplanes = 64
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = self._make_layer(block, 64, layers[0])
        self.layer2 = self._make_layer(block, 128, layers[1], stride=2)
        self.layer3 = self._make_layer(block, 256, layers[2], stride=2)
        self.layer4 = self._make_layer(block, 512, layers[3], stride=1, dilation=2)
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(512 * block.expansion, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")

    def _make_layer(self, block, planes, blocks, stride=1, dilation=1):
        down = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            down = nn.Sequential(conv1x1(self.inplanes, planes * block.expansion, stride),
                                 nn.BatchNorm2d(planes * block.expansion))
        seq = [block(self.inplanes, planes, stride, down)]           # first block: no dilation
        self.inplanes = planes * block.expansion
        seq += [block(self.inplanes, planes, dilation=dilation) for _ in range(1, blocks)]
        return nn.Sequential(*seq)


def resnet50(pretrained=False, **kwargs):
    return ResNet(Bottleneck, (3, 4, 6, 3), **kwargs)
