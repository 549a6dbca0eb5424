"""Parameter containers with the reference's factory names (/root/reference/compressai/layers/layers.py:29-43).

In this package an nn.Conv2d built by these factories only HOLDS the weights under the reference's
state_dict names; the arithmetic is done by csrc/conv.cu (tcgen05 implicit GEMM) driven from
compressai/models/_engine.py, which reads kernel size / stride / padding from the module."""
import torch.nn as nn


def conv3x3(in_ch, out_ch, stride=1):
    return nn.Conv2d(in_ch, out_ch, kernel_size=3, stride=stride, padding=1)


def subpel_conv3x3(in_ch, out_ch, r=1):
    return nn.Sequential(nn.Conv2d(in_ch, out_ch * r ** 2, kernel_size=3, padding=1), nn.PixelShuffle(r))


def conv1x1(in_ch, out_ch, stride=1):
    return nn.Conv2d(in_ch, out_ch, kernel_size=1, stride=stride)
