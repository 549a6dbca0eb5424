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

from ._engine import Engine
from .base import CompressionModel
from .utils import conv, update_registered_buffers

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


class SymmetricalTransFormer(CompressionModel):
    def __init__(self, pretrain_img_size=256, patch_size=2, in_chans=3, embed_dim=48, depths=[2, 2, 6, 2],
                 num_heads=[3, 6, 12, 24], window_size=4, num_slices=12, mlp_ratio=4., qkv_bias=True, qk_scale=None,
                 drop_rate=0., attn_drop_rate=0., drop_path_rate=0.2, norm_layer=nn.LayerNorm, patch_norm=True,
                 frozen_stages=-1, use_checkpoint=False):
        super().__init__()
        if (patch_size, in_chans, embed_dim, window_size, num_slices) != (2, 3, 48, 4, 12) or list(num_heads) != [3, 6, 12, 24]:
            raise NativeError("the CUDA STF path is built for the reference configuration "
                              "(patch 2, 3 input channels, embed_dim 48, window 4, heads 3/6/12/24, 12 slices)")
        self.num_layers = len(depths)
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

    # ---------------------------------------------------------------------------------- reference API
    def update(self, scale_table=None, force=False):
        if scale_table is None:
            scale_table = get_scale_table()
        updated = self.gaussian_conditional.update_scale_table(scale_table, force=force)
        updated |= super().update(force=force)
        return updated

    def load_state_dict(self, state_dict, strict=False):
        update_registered_buffers(self.gaussian_conditional, "gaussian_conditional",
                                  ["_quantized_cdf", "_offset", "_cdf_length", "scale_table"], state_dict)
        return super().load_state_dict(state_dict, strict=strict)

    @classmethod
    def from_state_dict(cls, state_dict):
        net = cls()
        net.load_state_dict(state_dict)
        return net

    # ---------------------------------------------------------------------------------- transforms
    def _check_input(self, x):
        if not x.is_cuda:
            raise NativeError("SymmetricalTransFormer runs on CUDA only (no CPU fallback); move the model and input to a CUDA device")
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError("expected an image batch [B, 3, H, W]")
        if x.shape[2] % 64 or x.shape[3] % 64:
            raise ValueError("H and W must be multiples of 64 (pad like the reference eval does, eval_model/__main__.py:103-115)")

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

    def _hyper_synthesis(self, z_hat_bf16, B, zh, zw):
        """h_mean_s / h_scale_s (stf.py:604-605) written straight into channels [0,384) of the support buffers."""
        e = self._engine
        P = 16 * zh * zw
        mean_sup = torch.zeros((B * P, 384 + 32 * 7), dtype=torch.bfloat16, device=z_hat_bf16.device)
        scale_sup = torch.zeros((B * P, 384 + 32 * 6), dtype=torch.bfloat16, device=z_hat_bf16.device)
        e.conv_stack(z_hat_bf16, B, zh, zw, self.h_mean_s, final_dtype=OUT_BF16, final_out=mean_sup)
        e.conv_stack(z_hat_bf16, B, zh, zw, self.h_scale_s, final_dtype=OUT_BF16, final_out=scale_sup)
        return mean_sup, scale_sup

    def _slice_loop(self, mode, B, h, w, mean_sup, scale_sup, y=None, decoder=None, y_lik=None):
        """The 12-slice channel-conditional loop (stf.py:611-631, 703-726, 754-776).

        mode "compress": returns (y_hat, symbols, indexes) with symbols/indexes int32 [B, 384*P] in stream order;
        mode "forward":  fills y_lik (NCHW fp32) and returns (y_hat, None, None);
        mode "decompress": symbols come from `decoder`."""
        e, gc = self._engine, self.gaussian_conditional
        dev = mean_sup.device
        P = h * w
        L = lib()
        st = stream_ptr()
        y_hat = torch.empty((B * P, 384), dtype=torch.float32, device=dev)
        table = gc.scale_table_device(dev)
        sym = idx = None
        if mode == "compress":
            sym = torch.empty((B, 384 * P), dtype=torch.int32, device=dev)
            idx = torch.empty_like(sym)
        elif mode == "decompress":
            gct = gc.device_tables()
            idx_s = torch.empty((B, 32 * P), dtype=torch.int32, device=dev)
            sym_s = torch.empty_like(idx_s)
        for i in range(self.num_slices):
            k = min(i, self.max_support_slices)
            mu, _, _ = e.conv_stack(mean_sup, B, h, w, self.cc_mean_transforms[i])
            sc, _, _ = e.conv_stack(scale_sup, B, h, w, self.cc_scale_transforms[i])
            v_mu, v_sc = view_bcp(mu, B, 32, P), view_bcp(sc, B, 32, P)
            v_hat = view_bcp(y_hat, B, 32, P, 32 * i)
            v_slot = view_bcp(mean_sup, B, 32, P, 384 + 32 * k)  # pre-LRP ŷ_i, input of lrp_transforms[i]
            if mode == "compress":
                check(L.icm_gc_quantize_index(view_bcp(y, B, 32, P, 32 * i), v_mu, v_sc, B, 32, P, table.data_ptr(), table.numel(),
                                              gc._scale_bound_f, sym.data_ptr(), idx.data_ptr(), 384 * P, 32 * P * i,
                                              v_hat, v_slot, NULL_VIEW, st), "icm_gc_quantize_index")
            elif mode == "forward":
                v_lik = View(y_lik.data_ptr() + 32 * i * P * 4, 384 * P, P, 1)
                check(L.icm_gc_likelihood(view_bcp(y, B, 32, P, 32 * i), v_mu, v_sc, B, 32, P, gc._scale_bound_f,
                                          gc.likelihood_bound if gc.use_likelihood_bound else 0.0, v_hat, v_lik, v_slot, NULL_VIEW, st),
                      "icm_gc_likelihood")
            else:
                check(L.icm_gc_build_indexes(v_sc, B, 32, P, table.data_ptr(), table.numel(), gc._scale_bound_f,
                                             idx_s.data_ptr(), 32 * P, 0, st), "icm_gc_build_indexes")
                decoder.decode_step(gct, idx_s, out=sym_s)
                check(L.icm_gc_dequantize(sym_s.data_ptr(), 32 * P, 0, v_mu, B, 32, P, v_hat, v_slot, NULL_VIEW, st), "icm_gc_dequantize")
            lrp, _, _ = e.conv_stack(mean_sup, B, h, w, self.lrp_transforms[i], final_act=ACT_HALF_TANH)
            keep = i < self.max_support_slices  # only the first six decoded slices are ever used as support (stf.py:612)
            check(L.icm_add_lrp(v_hat, view_bcp(lrp, B, 32, P), B, 32, P,
                                v_slot if keep else NULL_VIEW, view_bcp(scale_sup, B, 32, P, 384 + 32 * i) if keep else NULL_VIEW, st),
                  "icm_add_lrp")
        return y_hat, sym, idx

    def _hyper_analysis(self, y, B, h, w):
        e = self._engine
        z, zh, zw = e.conv_stack(e.cast_bf16(y), B, h, w, self.h_a)  # fp32 [B*zh*zw, 192]
        return z, zh, zw

    # ---------------------------------------------------------------------------------- forward / codec
    @torch.no_grad()
    def forward(self, x):
        """Eval-mode forward (stf.py:582-645): {"x_hat" (unclamped), "likelihoods": {"y", "z"}} in NCHW."""
        if self.training:
            raise NativeError("training-mode forward (noise quantisation + autograd) is not part of the CUDA inference path; call .eval()")
        self._check_input(x)
        eb = self.entropy_bottleneck
        B = x.shape[0]
        y, h, w = self._analysis(x)
        z, zh, zw = self._hyper_analysis(y, B, h, w)
        Pz, P = zh * zw, h * w
        z_hat = torch.empty((B * Pz, 192), dtype=torch.bfloat16, device=x.device)
        z_lik = torch.empty((B, 192, zh, zw), dtype=torch.float32, device=x.device)
        check(lib().icm_eb_process(1, view_bcp(z, B, 192, Pz), B, 192, Pz, eb.packed_params().data_ptr(),
                                   eb.likelihood_bound if eb.use_likelihood_bound else 0.0, None, None, NULL_VIEW,
                                   view_bcp(z_hat, B, 192, Pz), View(z_lik.data_ptr(), 192 * Pz, Pz, 1), stream_ptr()), "icm_eb_process")
        mean_sup, scale_sup = self._hyper_synthesis(z_hat, B, zh, zw)
        y_lik = torch.empty((B, 384, h, w), dtype=torch.float32, device=x.device)
        y_hat, _, _ = self._slice_loop("forward", B, h, w, mean_sup, scale_sup, y=y, y_lik=y_lik)
        x_hat = self._synthesis(y_hat, B, h, w, clamp=False)
        return {"x_hat": x_hat, "likelihoods": {"y": y_lik, "z": z_lik}}

    # ---------------------------------------------------------------------------------- micro-batching
    # The two rANS coders are latency-bound (one warp per image stream), the transforms are throughput-bound.
    # A batch is therefore split into `micro_batches` parts that run on their own CUDA streams, so that the
    # coder of one part overlaps the convolutions of another (the conv kernel's dynamic tile scheduler
    # tolerates SMs that are held by coder CTAs).  Results are identical to the unsplit run (kernels are batch-invariant).
    micro_batches = 2
    micro_batch_min = 16  # only split batches at least this large

    def _part_ranges(self, B):
        n = self.micro_batches if B >= self.micro_batch_min else 1
        n = max(1, min(n, B))
        base, extra = divmod(B, n)
        out, lo = [], 0
        for i in range(n):
            hi = lo + base + (1 if i < extra else 0)
            out.append((lo, hi))
            lo = hi
        return out

    def _run_parts(self, fns, coder_streams):
        """Run the callables on side streams (one each), joined back into the current stream."""
        if len(fns) == 1:
            return [fns[0]()]
        cur = torch.cuda.current_stream()
        pool = getattr(self, "_side_streams", None)
        if pool is None or len(pool) < len(fns) or pool[0].device != cur.device:
            pool = [torch.cuda.Stream(device=cur.device) for _ in fns]
            self._side_streams = pool
        outs = []
        for st, fn in zip(pool, fns):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                outs.append(fn())
        for st in pool[:len(fns)]:
            cur.wait_stream(st)
        return outs

    @staticmethod
    def _hand_over(t):
        """A tensor produced on a side stream is consumed on the current stream from now on."""
        if isinstance(t, torch.Tensor):
            t.record_stream(torch.cuda.current_stream())
        return t

    def _compress_part(self, x):
        """compress() of one micro-batch, fully asynchronous: device-resident streams."""
        eb = self.entropy_bottleneck
        B = x.shape[0]
        y, h, w = self._analysis(x)
        z, zh, zw = self._hyper_analysis(y, B, h, w)
        Pz = zh * zw
        z_sym = torch.empty((B, 192 * Pz), dtype=torch.int32, device=x.device)
        z_idx = torch.empty_like(z_sym)
        z_hat = torch.empty((B * Pz, 192), dtype=torch.bfloat16, device=x.device)
        check(lib().icm_eb_process(0, view_bcp(z, B, 192, Pz), B, 192, Pz, eb.packed_params().data_ptr(), 0.0,
                                   z_sym.data_ptr(), z_idx.data_ptr(), NULL_VIEW, view_bcp(z_hat, B, 192, Pz), NULL_VIEW, stream_ptr()),
              "icm_eb_process")
        mean_sup, scale_sup = self._hyper_synthesis(z_hat, B, zh, zw)
        _, sym, idx = self._slice_loop("compress", B, h, w, mean_sup, scale_sup, y=y)
        z_str = ans.encode_streams(eb.device_tables(), z_sym, z_idx, return_device="async")
        y_str = ans.encode_streams(self.gaussian_conditional.device_tables(), sym, idx, return_device="async")
        return {"y": y_str, "z": z_str, "shape": (zh, zw), "retry": (sym, idx, z_sym, z_idx)}

    @torch.no_grad()
    def compress(self, x, device_strings=False):
        """stf.py:671-732.  Returns {"strings": [[y_0..y_{B-1}], [z_0..z_{B-1}]], "shape": (h/4, w/4)}.

        device_strings=True keeps the streams on the GPU: "strings" is then [[(packed_y, sizes_y), ...], [(packed_z,
        sizes_z), ...]] with one (uint8, int32) CUDA tensor pair per micro-batch and no host synchronisation;
        decompress() accepts that form as is."""
        self._check_input(x)
        B = x.shape[0]
        parts = self._part_ranges(B)
        outs = self._run_parts([(lambda lo=lo, hi=hi: self._compress_part(x[lo:hi])) for lo, hi in parts], coder_streams=parts[0][1] - parts[0][0])
        shape = torch.Size(outs[0]["shape"])
        for o in outs:
            for k in ("y", "z"):
                self._hand_over(o[k][0]); self._hand_over(o[k][1])
        if device_strings:
            return {"strings": [[o["y"] for o in outs], [o["z"] for o in outs]], "shape": shape}
        gc_t, eb_t = self.gaussian_conditional.device_tables(), self.entropy_bottleneck.device_tables()
        y_strings, z_strings = [], []
        for o in outs:
            sym, idx, z_sym, z_idx = o["retry"]
            y_strings += ans.strings_to_host(*o["y"], retry=lambda: ans.encode_streams(gc_t, sym, idx))
            z_strings += ans.strings_to_host(*o["z"], retry=lambda: ans.encode_streams(eb_t, z_sym, z_idx))
        return {"strings": [y_strings, z_strings], "shape": shape}

    def _decompress_part(self, y_str, z_str, B, zh, zw, on_device, decoders=None):
        eb = self.entropy_bottleneck
        dev = eb.quantiles.device
        Pz = zh * zw
        h, w = 4 * zh, 4 * zw
        zdec = decoders[0] if decoders is not None else ans.acquire_decoder(B)
        zdec.set_streams_device(*z_str) if on_device else zdec.set_streams(z_str)
        z_idx = torch.arange(192, dtype=torch.int32, device=dev).repeat_interleave(Pz).repeat(B, 1)
        z_sym = zdec.decode_step(eb.device_tables(), z_idx)
        z_hat = torch.empty((B * Pz, 192), dtype=torch.bfloat16, device=dev)
        check(lib().icm_eb_process(2, NULL_VIEW, B, 192, Pz, eb.packed_params().data_ptr(), 0.0, z_sym.data_ptr(), None,
                                   NULL_VIEW, view_bcp(z_hat, B, 192, Pz), NULL_VIEW, stream_ptr()), "icm_eb_process")
        mean_sup, scale_sup = self._hyper_synthesis(z_hat, B, zh, zw)
        ydec = decoders[1] if decoders is not None else ans.acquire_decoder(B)
        ydec.set_streams_device(*y_str) if on_device else ydec.set_streams(y_str)
        y_hat, _, _ = self._slice_loop("decompress", B, h, w, mean_sup, scale_sup, decoder=ydec)
        x_hat = self._synthesis(y_hat, B, h, w, clamp=True)
        return x_hat, (zdec, ydec)

    @torch.no_grad()
    def decompress(self, strings, shape):
        """stf.py:734-785 for any batch size: strings = [[y strings], [z strings]], one of each per image
        (or the device-resident form produced by compress(..., device_strings=True))."""
        assert isinstance(strings, list) and len(strings) == 2
        eb = self.entropy_bottleneck
        dev = eb.quantiles.device
        if dev.type != "cuda":
            raise NativeError("SymmetricalTransFormer runs on CUDA only (no CPU fallback)")
        y_strings, z_strings = strings
        zh, zw = int(shape[0]), int(shape[1])
        on_device = len(z_strings) > 0 and isinstance(z_strings[0], tuple)
        if on_device:
            jobs = [(y, z, z[1].numel() - 1) for y, z in zip(y_strings, z_strings)]
        else:
            if len(y_strings) != len(z_strings):
                raise ValueError("need one y-string and one z-string per image")
            jobs = [(y_strings[lo:hi], z_strings[lo:hi], hi - lo) for lo, hi in self._part_ranges(len(z_strings))]
        outs = self._run_parts([(lambda y=y, z=z, n=n: self._decompress_part(y, z, n, zh, zw, on_device)) for y, z, n in jobs],
                               coder_streams=jobs[0][2])
        xs = [self._hand_over(o[0]) for o in outs]
        try:
            if not on_device:  # (status is a host read; the device-resident path leaves it to the caller)
                for _, decs in outs:
                    for d in decs:
                        d.check_status()
        finally:
            for _, decs in outs:
                for d in decs:
                    ans.release_decoder(d)
        return {"x_hat": xs[0] if len(xs) == 1 else torch.cat(xs, 0)}
