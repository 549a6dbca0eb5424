"""Window-attention parameter holders of WACNN (reference: /root/reference/compressai/layers/win_attention.py:37-207).
The computation is csrc/wacnn.cu:win_attention_kernel plus two tcgen05 linears (qkv, proj)."""
import torch
import torch.nn as nn


class WindowAttention(nn.Module):
    def __init__(self, dim=192, window_size=(8, 8), num_heads=8):
        super().__init__()
        self.dim, self.window_size, self.num_heads = dim, tuple(window_size), num_heads
        w = self.window_size[0]
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * w - 1) * (2 * w - 1), num_heads))
        t = torch.arange(w * w)
        th, tw = t // w, t % w
        self.register_buffer("relative_position_index", (th[:, None] - th[None, :] + w - 1) * (2 * w - 1) + (tw[:, None] - tw[None, :] + w - 1))
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02)


class WinBasedAttention(nn.Module):
    """x + proj(window_attention(qkv(x))) with optional cyclic shift; no norm, no MLP, no padding."""

    def __init__(self, dim=192, num_heads=8, window_size=8, shift_size=0):
        super().__init__()
        assert 0 <= shift_size < window_size, "shift_size must in 0-window_size"
        self.dim, self.num_heads, self.window_size, self.shift_size = dim, num_heads, window_size, shift_size
        self.attn = WindowAttention(dim, (window_size, window_size), num_heads)
