from .layers import conv1x1, conv3x3, subpel_conv3x3

__all__ = ["conv3x3", "subpel_conv3x3", "conv1x1"]
