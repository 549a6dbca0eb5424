"""Parameter containers with the reference's factory names (/root/reference/compressai/layers/layers.py:29-43).

In this package an nn.Conv2d built by these factories only HOLDS the weights under the reference's
state_dict names; the arithmetic is done by csrc/conv.cu (tcgen05 implicit GEMM) driven from
compressai/models/_engine.py, which reads kernel size / stride / padding from the module."""
import torch.nn as nn

from .win_attention import WinBasedAttention


def conv3x3(in_ch, out_ch, stride=1):
    return nn.Conv2d(in_ch, out_ch, kernel_size=3, stride=stride, padding=1)


def subpel_conv3x3(in_ch, out_ch, r=1):
    return nn.Sequential(nn.Conv2d(in_ch, out_ch * r ** 2, kernel_size=3, padding=1), nn.PixelShuffle(r))


def conv1x1(in_ch, out_ch, stride=1):
    return nn.Conv2d(in_ch, out_ch, kernel_size=1, stride=stride)


class ResidualUnit(nn.Module):
    """GELU(x + conv1x1(GELU(conv3x3(GELU(conv1x1(x))))))  (reference layers.py:52-75)."""

    def __init__(self, N):
        super().__init__()
        self.conv = nn.Sequential(conv1x1(N, N // 2), nn.GELU(), conv3x3(N // 2, N // 2), nn.GELU(), conv1x1(N // 2, N))
        self.relu = nn.GELU()


class Win_noShift_Attention(nn.Module):
    """out = x + conv_a(x) * sigmoid(conv_b(x))  (reference layers.py:45-89; shifted despite the name)."""

    def __init__(self, dim, num_heads=8, window_size=8, shift_size=0):
        super().__init__()
        self.conv_a = nn.Sequential(ResidualUnit(dim), ResidualUnit(dim), ResidualUnit(dim))
        self.conv_b = nn.Sequential(WinBasedAttention(dim=dim, num_heads=num_heads, window_size=window_size, shift_size=shift_size),
                                    ResidualUnit(dim), ResidualUnit(dim), ResidualUnit(dim), conv1x1(dim, dim))
