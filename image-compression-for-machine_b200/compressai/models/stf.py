"""SymmetricalTransFormer ("stf") on the B200-native kernels.

Drop-in for /root/reference/compressai/models/stf.py: same constructor defaults, the same parameter and
buffer names (so reference checkpoints load), and the same
forward / compress / decompress / update / load_state_dict / from_state_dict contract (stf.py:582-785).
The nn.Modules below only hold parameters; all arithmetic runs in the C-ABI CUDA library:

  patch embed, LayerNorm, window attention, 48->3 conv      csrc/transforms.cu
  every linear and 3x3 / 5x5 convolution                     csrc/conv.cu (tcgen05 + TMA implicit GEMM)
  quantise / indexes / likelihoods                           csrc/entropy.cu
  rANS                                                        csrc/rans.cu

Layout: activations are channels-last throughout (tokens [B, H*W, C] == NHWC), so none of the reference's
permute/contiguous/roll/window_partition/cat copies exist.  The context model's growing concatenations
(stf.py:613-626) are two persistent support buffers [B, h, w, 384 + 32*7] whose channel slots are written
in place by the producing kernels.

Batch semantics: `compress` returns ONE y-string PER IMAGE (strings[0][b]); each equals what the reference
returns for that image alone (the reference concatenates the batch into a single stream that its own
decompress cannot read back, stf.py:718-730,767 -- SURVEY.md §8b "Batch semantics").
"""
import math

import torch
import torch.nn as nn

from compressai import ans
from compressai._native import (ACT_HALF_TANH, NULL_VIEW, OUT_BF16, OUT_F32, NativeError, View, check, lib, stream_ptr,
                                view_bcp)
from compressai.entropy_models import EntropyBottleneck, GaussianConditional
from compressai.layers import conv3x3, subpel_conv3x3

from ._context import ChannelContextCodec
from ._engine import Engine
from .utils import conv

SCALES_MIN, SCALES_MAX, SCALES_LEVELS = 0.11, 256, 64


def get_scale_table(min=SCALES_MIN, max=SCALES_MAX, levels=SCALES_LEVELS):
    return torch.exp(torch.linspace(math.log(min), math.log(max), levels))


# ------------------------------------------------------------------------------------------------
# parameter holders (names follow stf.py:24-381)
class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.fc2 = nn.Linear(hidden_features, in_features)


class WindowAttention(nn.Module):
    def __init__(self, dim, window_size, num_heads):
        super().__init__()
        self.dim, self.window_size, self.num_heads = dim, window_size, num_heads
        w = window_size
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * w - 1) * (2 * w - 1), num_heads))
        t = torch.arange(w * w)
        th, tw = t // w, t % w
        index = (th[:, None] - th[None, :] + w - 1) * (2 * w - 1) + (tw[:, None] - tw[None, :] + w - 1)
        self.register_buffer("relative_position_index", index)  # kept for checkpoint compatibility; the kernel derives it
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02)


class SwinTransformerBlock(nn.Module):
    def __init__(self, dim, num_heads, window_size, shift_size, mlp_ratio):
        super().__init__()
        self.dim, self.window_size, self.shift_size = dim, window_size, shift_size
        self.norm1 = nn.LayerNorm(dim)
        self.attn = WindowAttention(dim, window_size, num_heads)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))


class PatchMerging(nn.Module):
    kind = "merge"

    def __init__(self, dim):
        super().__init__()
        self.reduction = nn.Linear(4 * dim, 2 * dim, bias=False)
        self.norm = nn.LayerNorm(4 * dim)


class PatchSplit(nn.Module):
    kind = "split"

    def __init__(self, dim):
        super().__init__()
        self.reduction = nn.Linear(dim, dim * 2, bias=False)
        self.norm = nn.LayerNorm(dim)


class BasicLayer(nn.Module):
    def __init__(self, dim, depth, num_heads, window_size, mlp_ratio, downsample):
        super().__init__()
        self.blocks = nn.ModuleList(
            SwinTransformerBlock(dim, num_heads, window_size, 0 if i % 2 == 0 else window_size // 2, mlp_ratio) for i in range(depth))
        self.downsample = downsample(dim) if downsample is not None else None


class PatchEmbed(nn.Module):
    def __init__(self, patch_size, in_chans, embed_dim):
        super().__init__()
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = nn.LayerNorm(embed_dim)


def _context_stack(cin):
    chans = (224, 176, 128, 64, 32)
    mods = []
    for i, c in enumerate(chans):
        mods.append(conv(cin, c, stride=1, kernel_size=3))
        if i < len(chans) - 1:
            mods.append(nn.GELU())
        cin = c
    return nn.Sequential(*mods)


class SymmetricalTransFormer(ChannelContextCodec):
    def __init__(self, pretrain_img_size=256, patch_size=2, in_chans=3, embed_dim=48, depths=[2, 2, 6, 2],
                 num_heads=[3, 6, 12, 24], window_size=4, num_slices=12, mlp_ratio=4., qkv_bias=True, qk_scale=None,
                 drop_rate=0., attn_drop_rate=0., drop_path_rate=0.2, norm_layer=nn.LayerNorm, patch_norm=True,
                 frozen_stages=-1, use_checkpoint=False):
        super().__init__()
        if (patch_size, in_chans, embed_dim, window_size, num_slices) != (2, 3, 48, 4, 12) or list(num_heads) != [3, 6, 12, 24]:
            raise NativeError("the CUDA STF path is built for the reference configuration "
                              "(patch 2, 3 input channels, embed_dim 48, window 4, heads 3/6/12/24, 12 slices)")
        self.num_layers = len(depths)
        self.latent_channels = embed_dim * 8
        self.hyper_channels = embed_dim * 4
        self.embed_dim = embed_dim
        self.num_slices = num_slices
        self.max_support_slices = num_slices // 2
        self.patch_embed = PatchEmbed(patch_size, in_chans, embed_dim)
        self.layers = nn.ModuleList(
            BasicLayer(embed_dim << i, depths[i], num_heads[i], window_size, mlp_ratio, PatchMerging if i < self.num_layers - 1 else None)
            for i in range(self.num_layers))
        rd, rh = depths[::-1], num_heads[::-1]
        self.syn_layers = nn.ModuleList(
            BasicLayer(embed_dim << (3 - i), rd[i], rh[i], window_size, mlp_ratio, PatchSplit if i < self.num_layers - 1 else None)
            for i in range(self.num_layers))
        self.end_conv = nn.Sequential(nn.Conv2d(embed_dim, embed_dim * patch_size ** 2, kernel_size=5, stride=1, padding=2),
                                      nn.PixelShuffle(patch_size), nn.Conv2d(embed_dim, 3, kernel_size=3, stride=1, padding=1))
        self.g_a = None
        self.g_s = None
        G = nn.GELU
        self.h_a = nn.Sequential(conv3x3(384, 384), G(), conv3x3(384, 336), G(), conv3x3(336, 288, stride=2), G(),
                                 conv3x3(288, 240), G(), conv3x3(240, 192, stride=2))
        hs = lambda: nn.Sequential(conv3x3(192, 240), G(), subpel_conv3x3(240, 288, 2), G(), conv3x3(288, 336), G(),
                                   subpel_conv3x3(336, 384, 2), G(), conv3x3(384, 384))
        self.h_mean_s = hs()
        self.h_scale_s = hs()
        self.cc_mean_transforms = nn.ModuleList(_context_stack(384 + 32 * min(i, 6)) for i in range(num_slices))
        self.cc_scale_transforms = nn.ModuleList(_context_stack(384 + 32 * min(i, 6)) for i in range(num_slices))
        self.lrp_transforms = nn.ModuleList(_context_stack(384 + 32 * min(i + 1, 7)) for i in range(num_slices))
        self.entropy_bottleneck = EntropyBottleneck(embed_dim * 4)
        self.gaussian_conditional = GaussianConditional(None)
        self._engine = Engine(self)

    scale_table_fn = staticmethod(get_scale_table)

    def _train_forward(self, x):
        from ._train import stf_train_forward

        return stf_train_forward(self, x, rng=self.train_rng, fused=self.train_fused)

    # ---------------------------------------------------------------------------------- transforms
    def _analysis(self, x):
        """g_a (stf.py:584-595): image [B,3,H,W] -> y fp32 channels-last [B*h*w, 384]."""
        e = self._engine
        B, _, H, W = x.shape
        x = x.detach().float().contiguous()
        pe = self.patch_embed
        H2, W2 = H // 2, W // 2
        t = torch.empty((B * H2 * W2, self.embed_dim), dtype=torch.float32, device=x.device)
        check(lib().icm_patch_embed(x.data_ptr(), e.f32(pe.proj.weight).data_ptr(), e.f32(pe.proj.bias).data_ptr(),
                                    e.f32(pe.norm.weight).data_ptr(), e.f32(pe.norm.bias).data_ptr(), t.data_ptr(),
                                    B, H, W, self.embed_dim, stream_ptr()), "icm_patch_embed")
        h, w = H2, W2
        for layer in self.layers:
            t, h, w = e.stage(t, B, h, w, layer)
        return t, h, w

    def _synthesis(self, y_hat, B, h, w, clamp):
        """g_s (stf.py:636-641, 778-784): y_hat fp32 channels-last [B*h*w, 384] (consumed) -> x_hat [B,3,16h,16w]."""
        e = self._engine
        t = y_hat
        for layer in self.syn_layers:
            t, h, w = e.stage(t, B, h, w, layer)
        tb = e.cast_bf16(t)
        u = e.conv(tb, B, h, w, e.packed(self.end_conv[0], ps=2))  # bf16 [B*2h*2w, 48]
        H, W = 2 * h, 2 * w
        out = torch.empty((B, 3, H, W), dtype=torch.float32, device=t.device)
        c2 = self.end_conv[2]
        check(lib().icm_final_conv(u.data_ptr(), e.f32(c2.weight).data_ptr(), e.f32(c2.bias).data_ptr(), out.data_ptr(),
                                   B, H, W, self.embed_dim, 1 if clamp else 0, stream_ptr()), "icm_final_conv")
        return out
