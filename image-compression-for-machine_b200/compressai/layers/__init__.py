from .gdn import GDN
from .layers import ResidualUnit, Win_noShift_Attention, conv1x1, conv3x3, subpel_conv3x3
from .win_attention import WinBasedAttention, WindowAttention

__all__ = ["GDN", "conv3x3", "subpel_conv3x3", "conv1x1", "Win_noShift_Attention", "WinBasedAttention", "WindowAttention", "ResidualUnit"]
