"""Inception-V3, CIFAR-10 variant (benchmark config #4), 21 638 954 parameters.

Architecture parity: reference ``models.py:96-392`` — the torchvision Inception-V3 body
(Mixed_5b .. Mixed_7c) behind a single stride-1 3x3 stem (``Conv2d_1a_3x3`` 3->192), global average
pool, dropout, ``fc`` 2048->num_classes; every conv is conv(bias=False)+BN(eps=1e-3)+ReLU.  Module and
parameter names match (``Mixed_6c.branch7x7dbl_4.conv.weight`` ...), so state dicts interchange.

Written table-driven: a Mixed block is a list of branches, a branch is a chain of conv specs with an
optional leading pool; blocks with forked tails (Mixed_7b/7c) list the fork explicitly.
"""
import torch
import torch.nn.functional as F
from torch import nn

from .. import nn as rnn


class BasicConv2d(nn.Module):
    def __init__(self, in_channels, out_channels, **kwargs):
        super().__init__()
        self.conv = rnn.Conv2d(in_channels, out_channels, bias=False, bn_stats=True, **kwargs)
        self.bn = rnn.BatchNorm2d(out_channels, eps=0.001, relu=True)      # BN + ReLU in one kernel

    def forward(self, x):
        return self.bn(self.conv(x))


def _k(kernel, **kw):
    return dict(kernel_size=kernel, **kw)


class Mixed(nn.Module):
    """Generic Inception block.  ``branches``: list of (pool, [(name, out_ch, conv kwargs)...], fork)
    where ``pool`` in {None,'avg','max'} is applied first ('max' with no convs = pass-through pool
    branch) and ``fork`` is an optional list of (name, out_ch, kwargs) applied in parallel to the
    chain's output and concatenated."""

    def __init__(self, in_channels, branches):
        super().__init__()
        self.program = []
        for pool, chain, fork in branches:
            ch = in_channels
            names = []
            for name, out_ch, kw in chain:
                self.add_module(name, BasicConv2d(ch, out_ch, **kw))
                names.append(name)
                ch = out_ch
            fork_names = []
            for name, out_ch, kw in fork or ():
                self.add_module(name, BasicConv2d(ch, out_ch, **kw))
                fork_names.append(name)
            self.program.append((pool, names, fork_names))
        self.avg3 = rnn.AvgPool2d(3, stride=1, padding=1)
        self.max3 = rnn.MaxPool2d(3, stride=2)

    def forward(self, x):
        outs = []
        for pool, names, fork_names in self.program:
            y = x
            if pool == 'avg':
                y = self.avg3(y)
            elif pool == 'max':
                y = self.max3(y)
            for n in names:
                y = getattr(self, n)(y)
            if fork_names:
                y = torch.cat([getattr(self, n)(y) for n in fork_names], 1)
            outs.append(y)
        return torch.cat(outs, 1)


def _inception_a(cin, pool_features):
    return Mixed(cin, [
        (None, [('branch1x1', 64, _k(1))], None),
        (None, [('branch5x5_1', 48, _k(1)), ('branch5x5_2', 64, _k(5, padding=2))], None),
        (None, [('branch3x3dbl_1', 64, _k(1)), ('branch3x3dbl_2', 96, _k(3, padding=1)), ('branch3x3dbl_3', 96, _k(3, padding=1))], None),
        ('avg', [('branch_pool', pool_features, _k(1))], None)])


def _inception_b(cin):
    return Mixed(cin, [
        (None, [('branch3x3', 384, _k(3, stride=2))], None),
        (None, [('branch3x3dbl_1', 64, _k(1)), ('branch3x3dbl_2', 96, _k(3, padding=1)), ('branch3x3dbl_3', 96, _k(3, stride=2))], None),
        ('max', [], None)])


def _inception_c(cin, c7):
    h, v = _k((1, 7), padding=(0, 3)), _k((7, 1), padding=(3, 0))
    return Mixed(cin, [
        (None, [('branch1x1', 192, _k(1))], None),
        (None, [('branch7x7_1', c7, _k(1)), ('branch7x7_2', c7, h), ('branch7x7_3', 192, v)], None),
        (None, [('branch7x7dbl_1', c7, _k(1)), ('branch7x7dbl_2', c7, v), ('branch7x7dbl_3', c7, h),
                ('branch7x7dbl_4', c7, v), ('branch7x7dbl_5', 192, h)], None),
        ('avg', [('branch_pool', 192, _k(1))], None)])


def _inception_d(cin):
    return Mixed(cin, [
        (None, [('branch3x3_1', 192, _k(1)), ('branch3x3_2', 320, _k(3, stride=2))], None),
        (None, [('branch7x7x3_1', 192, _k(1)), ('branch7x7x3_2', 192, _k((1, 7), padding=(0, 3))),
                ('branch7x7x3_3', 192, _k((7, 1), padding=(3, 0))), ('branch7x7x3_4', 192, _k(3, stride=2))], None),
        ('max', [], None)])


def _inception_e(cin):
    h, v = _k((1, 3), padding=(0, 1)), _k((3, 1), padding=(1, 0))
    return Mixed(cin, [
        (None, [('branch1x1', 320, _k(1))], None),
        (None, [('branch3x3_1', 384, _k(1))], [('branch3x3_2a', 384, h), ('branch3x3_2b', 384, v)]),
        (None, [('branch3x3dbl_1', 448, _k(1)), ('branch3x3dbl_2', 384, _k(3, padding=1))],
         [('branch3x3dbl_3a', 384, h), ('branch3x3dbl_3b', 384, v)]),
        ('avg', [('branch_pool', 192, _k(1))], None)])


class Inception3(nn.Module):
    """Input ``[N, 3, 32, 32]`` -> logits ``[N, num_classes]``."""

    def __init__(self, num_classes=10, aux_logits=False, transform_input=False):
        super().__init__()
        assert not aux_logits, 'the CIFAR variant trains without the auxiliary head (reference models.py:98)'
        self.Conv2d_1a_3x3 = BasicConv2d(3, 192, kernel_size=3, stride=1, padding=1)
        self.Mixed_5b = _inception_a(192, 32)
        self.Mixed_5c = _inception_a(256, 64)
        self.Mixed_5d = _inception_a(288, 64)
        self.Mixed_6a = _inception_b(288)
        self.Mixed_6b = _inception_c(768, 128)
        self.Mixed_6c = _inception_c(768, 160)
        self.Mixed_6d = _inception_c(768, 160)
        self.Mixed_6e = _inception_c(768, 192)
        self.Mixed_7a = _inception_d(768)
        self.Mixed_7b = _inception_e(1280)
        self.Mixed_7c = _inception_e(2048)
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.dropout = rnn.Dropout(0.5)
        self.flatten = nn.Flatten(1)
        self.fc = nn.Linear(2048, num_classes)
        self._blocks = ['Conv2d_1a_3x3', 'Mixed_5b', 'Mixed_5c', 'Mixed_5d', 'Mixed_6a', 'Mixed_6b', 'Mixed_6c',
                        'Mixed_6d', 'Mixed_6e', 'Mixed_7a', 'Mixed_7b', 'Mixed_7c', 'avgpool', 'dropout', 'flatten', 'fc']

    def forward(self, x):
        for name in self._blocks:
            x = getattr(self, name)(x)
        return x


def inception_v3(**kwargs):
    return Inception3(**kwargs)
