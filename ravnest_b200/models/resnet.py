"""ResNet-50 (benchmark config #2).

Architecture / parameter-name parity with ``torchvision.models.resnet50`` which the reference uses
(``cluster_formation.py:23-25``, ``examples/resnet50/provider.py``): 7x7/2 stem, max-pool, bottleneck
stages [3, 4, 6, 3] with stride on the 3x3 conv, global average pool, ``fc``.  ``num_classes=200`` gives
the 23 917 832 parameters of the Tiny-ImageNet config.  torchvision state dicts load unchanged.
"""
import torch
from torch import nn

from .. import nn as rnn


class Bottleneck(nn.Module):
    """conv1x1 -> BN+ReLU -> conv3x3 -> BN+ReLU -> conv1x1 -> BN + identity + ReLU; each "BN(+add)+ReLU" is ONE fused kernel
    (rnn.BatchNorm2d).  Attribute / parameter names match torchvision so its checkpoints load."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=False):
        super().__init__()
        out = planes * self.expansion
        # identity blocks: conv1 also hands out the skip branch, so its data gradient lands IN the skip gradient (no add kernel)
        self.conv1 = rnn.Conv2d(inplanes, planes, 1, bias=False, fork=not downsample, bn_stats=True)
        self.bn1 = rnn.BatchNorm2d(planes, relu=True)
        self.conv2 = rnn.Conv2d(planes, planes, 3, stride=stride, padding=1, bias=False, bn_stats=True)
        self.bn2 = rnn.BatchNorm2d(planes, relu=True)
        self.conv3 = rnn.Conv2d(planes, out, 1, bias=False, bn_stats=True)
        self.bn3 = rnn.BatchNorm2d(out, relu=True)           # relu(BN(conv3) + identity)
        self.downsample = None
        if downsample:
            self.downsample = nn.Sequential(rnn.Conv2d(inplanes, out, 1, stride=stride, bias=False, bn_stats=True), rnn.BatchNorm2d(out))

    def forward(self, x):
        if getattr(self.conv1, 'fork', False):
            both = self.conv1(x)
            y, identity = both[0], both[1]
        else:
            y, identity = self.conv1(x), self.downsample(x)
        y = self.bn2(self.conv2(self.bn1(y)))
        return self.bn3(self.conv3(y), identity)


class ResNet(nn.Module):
    def __init__(self, layers=(3, 4, 6, 3), num_classes=1000, zero_init_residual=False):
        super().__init__()
        self.conv1 = rnn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False, bn_stats=True)
        self.bn1 = rnn.BatchNorm2d(64, relu=True)
        self.maxpool = rnn.MaxPool2d(3, stride=2, padding=1)
        inplanes = 64
        for i, (planes, blocks) in enumerate(zip((64, 128, 256, 512), layers)):
            stride = 1 if i == 0 else 2
            stage = [Bottleneck(inplanes, planes, stride, downsample=True)]
            inplanes = planes * Bottleneck.expansion
            stage += [Bottleneck(inplanes, planes) for _ in range(blocks - 1)]
            setattr(self, 'layer{}'.format(i + 1), nn.Sequential(*stage))
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.flatten = nn.Flatten(1)
        self.fc = nn.Linear(inplanes, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
        if zero_init_residual:
            for m in self.modules():
                if isinstance(m, Bottleneck):
                    nn.init.zeros_(m.bn3.weight)

    def forward(self, x):
        x = self.maxpool(self.bn1(self.conv1(x)))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.fc(self.flatten(self.avgpool(x)))


def resnet50(num_classes=1000, **kwargs):
    return ResNet((3, 4, 6, 3), num_classes=num_classes, **kwargs)
