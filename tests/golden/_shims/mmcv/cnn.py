import torch.nn as nn


class ConvModule(nn.Module):
    """mmcv.cnn.ConvModule without norm: children named .conv / .activate (state-dict compatible)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, norm_cfg=None, act_cfg=None):
        super().__init__()
        assert norm_cfg is None
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding)
        self.activate = nn.ReLU(inplace=True) if act_cfg is not None else None

    def forward(self, x):
        x = self.conv(x)
        return self.activate(x) if self.activate is not None else x
