"""The 16 426-parameter digits CNN used by plumbing config #1.

Architecture parity: reference ``models.py:3-44`` (conv3x3(1->16) - ReLU - pool - drop - BN - pool -
conv3x3(16->32) - ReLU - pool - drop - BN - flatten - FC(32->256) - ReLU - drop - BN1d - FC(256->10) - softmax).
Layer names are kept because ring ids / param-address maps are expressed in parameter names.
"""
from collections import OrderedDict

from torch import nn

from .. import nn as rnn


def _layers():
    return OrderedDict([
        ('conv2d_1', nn.Conv2d(1, 16, kernel_size=3, padding='same')),
        ('act_1', nn.ReLU()),
        ('maxpool2d_1', nn.MaxPool2d(2, stride=2)),
        ('drp_1', rnn.Dropout(0.25)),
        ('bn_1', nn.BatchNorm2d(16)),
        ('maxpool2d_2', nn.MaxPool2d(2, stride=2)),
        ('conv2d_2', nn.Conv2d(16, 32, kernel_size=3, padding='same')),
        ('act_2', nn.ReLU()),
        ('maxpool2d_3', nn.MaxPool2d(2, stride=2)),
        ('drp_2', rnn.Dropout(0.25)),
        ('bn_2', nn.BatchNorm2d(32)),
        ('flatten', nn.Flatten()),
        ('dense_1', nn.Linear(32, 256)),
        ('act_3', nn.ReLU()),
        ('drp_3', rnn.Dropout(0.4)),
        ('bn_3', nn.BatchNorm1d(256)),
        ('dense_2', nn.Linear(256, 10)),
        ('act_4', nn.Softmax(dim=-1)),
    ])


class CNN_Net(nn.Module):
    """Input ``[N, 1, 8, 8]`` -> class probabilities ``[N, 10]``."""

    def __init__(self):
        super().__init__()
        self._order = []
        for name, layer in _layers().items():
            self.add_module(name, layer)
            self._order.append(name)

    def forward(self, x):
        for name in self._order:
            x = getattr(self, name)(x)
        return x
