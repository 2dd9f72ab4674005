"""mmcv.ops stand-in: parameters as mmcv-full 1.x, arithmetic by torchvision.ops.deform_conv2d
(same MSRA DCNv2 lineage, channel layouts identical)."""
import math

import torch
import torch.nn as nn
import torchvision
from torch.nn.modules.utils import _pair


def modulated_deform_conv2d(input, offset, mask, weight, bias, stride, padding, dilation, groups, deform_groups):
    return torchvision.ops.deform_conv2d(input, offset, weight, bias, stride=stride, padding=padding,
                                         dilation=dilation, mask=mask)


class ModulatedDeformConv2d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 deform_groups=1, bias=True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = _pair(kernel_size)
        self.stride, self.padding, self.dilation = _pair(stride), _pair(padding), _pair(dilation)
        self.groups, self.deform_groups = groups, deform_groups
        self.weight = nn.Parameter(torch.Tensor(out_channels, in_channels // groups, *self.kernel_size))
        self.bias = nn.Parameter(torch.Tensor(out_channels)) if bias else None
        stdv = 1. / math.sqrt(in_channels * self.kernel_size[0] * self.kernel_size[1])
        self.weight.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.zero_()

    def forward(self, x, offset, mask):
        return modulated_deform_conv2d(x, offset, mask, self.weight, self.bias, self.stride, self.padding,
                                       self.dilation, self.groups, self.deform_groups)
