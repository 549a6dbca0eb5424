"""Host-side driver of the CUDA transform kernels: weight packing and thin call wrappers.

Activations are channels-last.  `conv()` runs csrc/conv.cu (tcgen05 implicit GEMM) for nn.Conv2d / nn.Linear
parameter holders; `swin_block()` etc. string the kernels of csrc/transforms.cu together.  PyTorch supplies
device memory and the stream; no torch operator does arithmetic on this path.
"""
import ctypes as C
import os

import torch

from compressai import _native
from compressai._native import (ACT_GELU, ACT_NONE, ACT_RSQRT, ACT_SIGMOID, ACT_SQRT, OUT_BF16, OUT_F32, RES_ADD, RES_ADD_BEFORE_ACT,
                                RES_MUL, ConvArgs, ConvGroups, NativeError, check, lib, stream_ptr)

_PITCH64 = os.environ.get("ICM_PITCH64", "1") != "0"  # A/B switch for Engine._pitch64


class PackedConv:
    """bf16 [Cout_pad, KH*KW*Cin_pad] weight + fp32 bias on the device (icm_pack_conv_weight)."""

    __slots__ = ("w", "bias", "Cin", "Cout", "KH", "KW", "stride", "pad", "ps")

    def __init__(self, weight, bias, stride=1, pad=0, ps=0):
        w = weight.detach().float().contiguous()
        if w.dim() == 2:
            w = w[:, :, None, None]
        self.Cout, self.Cin, self.KH, self.KW = w.shape
        self.stride, self.pad, self.ps = int(stride), int(pad), int(ps)
        cin_pad = (self.Cin + 63) // 64 * 64
        cout_pad = (self.Cout + 15) // 16 * 16
        self.w = torch.empty((cout_pad, self.KH * self.KW * cin_pad), dtype=torch.bfloat16, device=w.device)
        check(lib().icm_pack_conv_weight(w.data_ptr(), self.Cout, self.Cin, self.KH, self.KW, cin_pad, cout_pad, self.ps,
                                         self.w.data_ptr(), stream_ptr()), "icm_pack_conv_weight")
        self.bias = None
        if bias is not None:
            b = bias.detach().float()
            if self.ps:  # GEMM column (i*r+j)*Cq + c holds conv channel c*r^2 + (i*r+j): permute the bias alike
                b = b.reshape(self.Cout // (self.ps * self.ps), self.ps * self.ps).t()
            self.bias = b.contiguous().reshape(-1).clone()


class PackedDeconv:
    """ConvTranspose2d(k5, s2, p2, output_padding 1) as a 3x3 convolution with 4*Cq phase-major output channels
    and the PixelShuffle(2) store (icm_pack_deconv_weight); Cq = Cout rounded up to 16."""

    __slots__ = ("w", "bias", "Cin", "Cout", "KH", "KW", "stride", "pad", "ps", "true_cout")

    def __init__(self, weight, bias):
        w = weight.detach().float().contiguous()  # [Cin, Cout, 5, 5]
        assert w.shape[2:] == (5, 5)
        self.Cin, self.true_cout = w.shape[0], w.shape[1]
        cq = (self.true_cout + 15) // 16 * 16
        cin_pad = (self.Cin + 63) // 64 * 64
        self.Cout, self.KH, self.KW, self.stride, self.pad, self.ps = 4 * cq, 3, 3, 1, 1, 2
        self.w = torch.empty((4 * cq, 9 * cin_pad), dtype=torch.bfloat16, device=w.device)
        check(lib().icm_pack_deconv_weight(w.data_ptr(), self.Cin, self.true_cout, cin_pad, cq, self.w.data_ptr(), stream_ptr()),
              "icm_pack_deconv_weight")
        b = torch.zeros(4, cq, dtype=torch.float32, device=w.device)
        if bias is not None:
            b[:, : self.true_cout] = bias.detach().float()
        self.bias = b.reshape(-1).contiguous()


class PackedGroup:
    """The packed weights of G same-shaped layers stacked [G * Cout_pad, K] (+ biases [G, Cout]) for icm_conv2d_grouped."""

    __slots__ = ("w", "bias", "Cin", "Cout", "KH", "KW", "stride", "pad", "ps", "G", "rows")

    def __init__(self, pks):
        p0 = pks[0]
        for pk in pks:
            assert (pk.Cin, pk.Cout, pk.KH, pk.KW, pk.stride, pk.pad, pk.ps) == (p0.Cin, p0.Cout, p0.KH, p0.KW, p0.stride, p0.pad, p0.ps)
            assert (pk.bias is None) == (p0.bias is None)
        self.Cin, self.Cout, self.KH, self.KW, self.stride, self.pad, self.ps = p0.Cin, p0.Cout, p0.KH, p0.KW, p0.stride, p0.pad, p0.ps
        self.G, self.rows = len(pks), p0.w.shape[0]
        self.w = torch.cat([pk.w for pk in pks], 0).contiguous()
        self.bias = torch.stack([pk.bias for pk in pks], 0).contiguous() if p0.bias is not None else None


class Engine:
    def __init__(self, model):
        self.model = model
        self._packed = {}
        self.warm = False  # True once every weight has been packed (first full run): micro-batching may start

    def invalidate(self):
        self._packed.clear()
        self.warm = False

    # ---------------------------------------------------------------------------------- weights
    @staticmethod
    def _versions(module):
        """(data_ptr, _version) of every parameter of a holder: in-place edits (optimizer steps, weight.mul_) bump _version,
        re-assignment changes data_ptr -- either invalidates the packed copy, like EntropyBottleneck.packed_params()."""
        return tuple((p.data_ptr(), p._version) for p in module._parameters.values() if p is not None)

    def packed(self, module, ps=0):
        key = (id(module), ps)
        hit = self._packed.get(key)
        pk = hit[0] if hit is not None and hit[2] == self._versions(module) else None
        if pk is None:
            if isinstance(module, torch.nn.Conv2d):
                pk = PackedConv(module.weight, module.bias, module.stride[0], module.padding[0], ps)
            elif isinstance(module, torch.nn.Linear):
                pk = PackedConv(module.weight, module.bias, 1, 0, ps)
            elif isinstance(module, torch.nn.ConvTranspose2d):
                pk = PackedDeconv(module.weight, module.bias)
            elif hasattr(module, "effective"):  # GDN: 1x1 GEMM with gamma', bias beta'
                gamma, beta = module.effective()
                pk = PackedConv(gamma, beta, 1, 0, 0)
            else:
                raise TypeError(type(module))
            self._packed[key] = (pk, module, self._versions(module))  # holding the module keeps its id() from being reused
            torch.cuda.current_stream().synchronize()  # packed before any other stream may use it (micro-batches)
        return pk

    def packed_group(self, modules, ps=0):
        """Stacked packed weights of same-shaped layers (one icm_conv2d_grouped launch)."""
        key = (tuple(id(m) for m in modules), ps, "group")
        vers = tuple(self._versions(m) for m in modules)
        hit = self._packed.get(key)
        if hit is not None and hit[2] == vers:
            return hit[0]
        pg = PackedGroup([self.packed(m, ps) for m in modules])
        self._packed[key] = (pg, tuple(modules), vers)
        torch.cuda.current_stream().synchronize()
        return pg

    def f32(self, p):
        key = (id(p), "f32")
        hit = self._packed.get(key)
        if hit is None or hit[2] != (p.data_ptr(), p._version):
            hit = (p.detach().float().contiguous(), p, (p.data_ptr(), p._version))
            self._packed[key] = hit
            torch.cuda.current_stream().synchronize()
        return hit[0]

    # ---------------------------------------------------------------------------------- kernels
    @staticmethod
    def _pitch64(c):
        """Row pitch (in channels) of an activation that only a following convolution reads: the next multiple of 64 channels
        = 128 bytes.  The TMA engine of an SM moves 64-channel boxes at 111 B/clk when every pixel's 128-byte piece is one
        aligned L2 line, 85 at a pitch of 448 bytes (C = 224) and 73 at 352 (C = 176) (tools/umma_issue_bench.cu), and in the
        N <= 176 layers of the context-model stacks that ingest takes as long as the MMAs: 176 -> 128 layers
        +6-9 %, 224 -> 176 +2.5 % (profiles/r02_conv_pitch_ab.txt)."""
        return (c + 63) // 64 * 64 if _PITCH64 else c

    def conv(self, x, B, H, W, pk, out=None, out_offset=0, act=ACT_NONE, out_dtype=OUT_BF16, residual=None, cin=None, res_mode=0,
             pitch64=False):
        """x: bf16 [.., pitch] channels-last holding B*H*W pixels.  Returns the output tensor
        ([B*Ho*Wo(*r*r), Cout(/r^2)] unless `out` is given, then writes channels [out_offset, +Cout) of it; with pitch64 the
        rows of a new output are padded to a multiple of 64 channels -- the pad is never written and never read: the next
        convolution reads Cin channels and TMA zero-fills the rest of its last 64-channel box)."""
        assert x.dtype == torch.bfloat16 and x.is_cuda
        in_pitch = x.shape[-1]
        cin = pk.Cin if cin is None else cin
        if cin != pk.Cin:
            raise NativeError(f"conv: input channels {cin} != weight channels {pk.Cin}")
        Ho = (H + 2 * pk.pad - pk.KH) // pk.stride + 1
        Wo = (W + 2 * pk.pad - pk.KW) // pk.stride + 1
        r = pk.ps if pk.ps else 1
        cout_eff = pk.Cout // (r * r)
        if out is None:
            out = torch.empty((B * Ho * r * Wo * r, self._pitch64(cout_eff) if pitch64 else cout_eff),
                              dtype=torch.float32 if out_dtype == OUT_F32 else torch.bfloat16, device=x.device)
        a = ConvArgs()
        a.inp, a.weight = x.data_ptr(), pk.w.data_ptr()
        a.bias = pk.bias.data_ptr() if pk.bias is not None else None
        a.out = out.data_ptr() + out_offset * out.element_size()
        a.residual = residual.data_ptr() if residual is not None else None
        a.B, a.H, a.W, a.Cin, a.in_pitch = B, H, W, cin, in_pitch
        a.Cout, a.out_pitch = pk.Cout, out.shape[-1]
        a.KH, a.KW, a.stride, a.pad = pk.KH, pk.KW, pk.stride, pk.pad
        a.act, a.out_dtype, a.pixel_shuffle = act, out_dtype, pk.ps
        a.res_pitch = residual.shape[-1] if residual is not None else 0
        a.res_dtype = OUT_BF16 if (residual is not None and residual.dtype == torch.bfloat16) else OUT_F32
        a.res_mode = res_mode
        check(lib().icm_conv2d(C.byref(a), stream_ptr()), "icm_conv2d")
        return out

    def conv_group(self, x, B, H, W, pg, in_images, in_offsets, tail_channels=None, out=None, out_offset=0, out_group_stride=None,
                   act=ACT_NONE, out_dtype=OUT_BF16, cin=None, pitch64=False):
        """G same-shaped convolutions in one launch (icm_conv2d_grouped).  x: bf16 channels-last tensor of `in_images` images;
        group g reads images [in_offsets[g], +B).  Output: stacked [G, B*Ho*Wo(*r*r), Cout(/r^2)] unless `out` is given, then group g
        writes at element offset out_offset + g * out_group_stride of it (row pitch = out.shape[-1])."""
        assert x.dtype == torch.bfloat16 and x.is_cuda
        cin = pg.Cin if cin is None else cin
        if cin != pg.Cin:
            raise NativeError(f"conv_group: input channels {cin} != weight channels {pg.Cin}")
        Ho = (H + 2 * pg.pad - pg.KH) // pg.stride + 1
        Wo = (W + 2 * pg.pad - pg.KW) // pg.stride + 1
        r = pg.ps if pg.ps else 1
        cout_eff = pg.Cout // (r * r)
        G = pg.G
        if out is None:
            out = torch.empty((G, B * Ho * r * Wo * r, self._pitch64(cout_eff) if pitch64 else cout_eff),
                              dtype=torch.float32 if out_dtype == OUT_F32 else torch.bfloat16, device=x.device)
            out_group_stride = out.shape[1] * out.shape[2]
        a = ConvArgs()
        a.inp, a.weight = x.data_ptr(), pg.w.data_ptr()
        a.bias = pg.bias.data_ptr() if pg.bias is not None else None
        a.out = out.data_ptr() + out_offset * out.element_size()
        a.residual = None
        a.B, a.H, a.W, a.Cin, a.in_pitch = B, H, W, cin, x.shape[-1]
        a.Cout, a.out_pitch = pg.Cout, out.shape[-1]
        a.KH, a.KW, a.stride, a.pad = pg.KH, pg.KW, pg.stride, pg.pad
        a.act, a.out_dtype, a.pixel_shuffle = act, out_dtype, pg.ps
        a.res_pitch, a.res_dtype, a.res_mode = 0, OUT_F32, 0
        g = ConvGroups()
        g.groups, g.in_images = G, in_images
        g.weight_group_rows = pg.rows
        g.bias_group_stride = pg.bias.shape[1] if pg.bias is not None else 0
        g.out_group_stride = out_group_stride
        for k in range(G):
            g.in_image_offset[k] = in_offsets[k]
            g.tail_channel[k] = tail_channels[k] if tail_channels is not None else -1
        check(lib().icm_conv2d_grouped(C.byref(a), C.byref(g), stream_ptr()), "icm_conv2d_grouped")
        return out

    def conv_stack_group(self, x, B, H, W, seqs, in_images, in_offsets, tail_channels=None, final_dtype=OUT_F32, final_act=ACT_NONE,
                         final_out=None, final_out_offset=0, final_out_group_stride=None, cin=None):
        """G same-shaped nn.Sequential conv stacks, layer by layer in grouped launches; the intermediate activations are
        stacked [G*B, H, W, C]."""
        per = [[(m, 0) if isinstance(m, torch.nn.Conv2d) else (m[0], m[1].upscale_factor)
                for m in seq if isinstance(m, (torch.nn.Conv2d, torch.nn.Sequential))] for seq in seqs]
        G, n = len(seqs), len(per[0])
        for k in range(n):
            last = k == n - 1
            ps = per[0][k][1]
            pg = self.packed_group([per[g][k][0] for g in range(G)], ps)
            kw = dict(cin=cin) if k == 0 else {}
            if k == 0:
                ii, io, tc = in_images, in_offsets, tail_channels
            else:
                ii, io, tc = G * B, [g * B for g in range(G)], None
            if last:
                x = self.conv_group(x, B, H, W, pg, ii, io, tc, out=final_out, out_offset=final_out_offset,
                                    out_group_stride=final_out_group_stride, act=final_act, out_dtype=final_dtype, **kw)
            else:
                x = self.conv_group(x, B, H, W, pg, ii, io, tc, act=ACT_GELU, pitch64=True, **kw)
            H = (H + 2 * pg.pad - pg.KH) // pg.stride + 1
            W = (W + 2 * pg.pad - pg.KW) // pg.stride + 1
            if ps:
                H, W = H * ps, W * ps
        return x, H, W

    def linear(self, x, pk, **kw):
        """x: bf16 [M, K] -> [M, N]."""
        return self.conv(x, 1, 1, x.shape[0], pk, **kw)

    def layernorm(self, x, norm, out_dtype=OUT_BF16, gather=None):
        """x fp32 [rows, C]; gather=(B,H,W): PatchMerging gather, output [B*ceil(H/2)*ceil(W/2), 4C]."""
        Cc = norm.normalized_shape[0]
        if gather is None:
            rows, B, H, W, g = x.shape[0], 0, 0, 0, 0
        else:
            B, H, W = gather
            rows, g = B * ((H + 1) // 2) * ((W + 1) // 2), 1
        out = torch.empty((rows, Cc), dtype=torch.bfloat16 if out_dtype == OUT_BF16 else torch.float32, device=x.device)
        check(lib().icm_layernorm(x.data_ptr(), self.f32(norm.weight).data_ptr(), self.f32(norm.bias).data_ptr(), out.data_ptr(),
                                  out_dtype, rows, Cc, g, B, H, W, stream_ptr()), "icm_layernorm")
        return out

    def cast_bf16(self, x):
        rows, Cc = x.shape
        out = torch.empty((rows, Cc), dtype=torch.bfloat16, device=x.device)
        check(lib().icm_cast_bf16(x.data_ptr(), rows, Cc, Cc, out.data_ptr(), Cc, stream_ptr()), "icm_cast_bf16")
        return out

    def window_attention(self, qkv, B, H, W, Cc, heads, window, shift, bias_table):
        out = torch.empty((B * H * W, Cc), dtype=torch.bfloat16, device=qkv.device)
        check(lib().icm_window_attention(qkv.data_ptr(), out.data_ptr(), self.f32(bias_table).data_ptr(), B, H, W, Cc, heads,
                                         window, shift, stream_ptr()), "icm_window_attention")
        return out

    # ---------------------------------------------------------------------------------- Swin pieces
    fused_block = True  # icm_swin_block for C in (48, 96): one pass over the residual stream instead of six launches

    def swin_block(self, x, B, H, W, blk, shifted):
        """x: fp32 tokens [B*H*W, C], updated in place (stf.py:149-199)."""
        Cc = x.shape[1]
        if self.fused_block and Cc in (48, 96) and blk.window_size == 4 and x.is_contiguous() and H % 4 == 0 and W % 4 == 0:
            at, mlp = blk.attn, blk.mlp
            qkv, proj, f1, f2 = self.packed(at.qkv), self.packed(at.proj), self.packed(mlp.fc1), self.packed(mlp.fc2)
            ptr = lambda t: t.data_ptr() if t is not None else None
            check(lib().icm_swin_block(x.data_ptr(), B, H, W, Cc, at.num_heads, blk.window_size, blk.window_size // 2 if shifted else 0, 3,
                                       qkv.w.data_ptr(), ptr(qkv.bias), proj.w.data_ptr(), ptr(proj.bias),
                                       self.f32(at.relative_position_bias_table).data_ptr(),
                                       self.f32(blk.norm1.weight).data_ptr(), self.f32(blk.norm1.bias).data_ptr(),
                                       f1.w.data_ptr(), ptr(f1.bias), f2.w.data_ptr(), ptr(f2.bias),
                                       self.f32(blk.norm2.weight).data_ptr(), self.f32(blk.norm2.bias).data_ptr(), stream_ptr()), "icm_swin_block")
            return x
        xn = self.layernorm(x, blk.norm1)
        qkv = self.linear(xn, self.packed(blk.attn.qkv))
        ao = self.window_attention(qkv, B, H, W, Cc, blk.attn.num_heads, blk.window_size, blk.window_size // 2 if shifted else 0,
                                   blk.attn.relative_position_bias_table)
        self.linear(ao, self.packed(blk.attn.proj), out=x, out_dtype=OUT_F32, residual=x)
        xn = self.layernorm(x, blk.norm2)
        self.mlp(xn, x, blk.mlp)
        return x

    fused_mlp = True  # icm_swin_mlp for C in (48, 96, 192): bit-identical to the two launches, without the hidden round trip

    def mlp(self, xn, x, mlp):
        """x += fc2(GELU(fc1(xn)))  (stf.py:34-40 with the residual of :198); x fp32 [M, C] in place, xn bf16 [M, C]."""
        Cc = x.shape[1]
        f1, f2 = self.packed(mlp.fc1), self.packed(mlp.fc2)
        if self.fused_mlp and Cc in (48, 96, 192) and x.is_contiguous() and xn.is_contiguous():
            check(lib().icm_swin_mlp(xn.data_ptr(), f1.w.data_ptr(), f1.bias.data_ptr() if f1.bias is not None else None, f2.w.data_ptr(),
                                     f2.bias.data_ptr() if f2.bias is not None else None, x.data_ptr(), x.shape[0], Cc, stream_ptr()), "icm_swin_mlp")
            return x
        h = self.linear(xn, f1, act=ACT_GELU)
        self.linear(h, f2, out=x, out_dtype=OUT_F32, residual=x)
        return x

    def stage(self, x, B, H, W, layer):
        for i, blk in enumerate(layer.blocks):
            x = self.swin_block(x, B, H, W, blk, i % 2 == 1)
        ds = layer.downsample
        if ds is None:
            return x, H, W
        if ds.kind == "merge":
            xn = self.layernorm(x, ds.norm, gather=(B, H, W))
            x = self.linear(xn, self.packed(ds.reduction), out_dtype=OUT_F32)
            return x, (H + 1) // 2, (W + 1) // 2
        xn = self.layernorm(x, ds.norm)
        x = self.conv(xn, B, H, W, self.packed(ds.reduction, ps=2), out_dtype=OUT_F32)
        return x, 2 * H, 2 * W

    def conv_stack(self, x, B, H, W, seq, final_dtype=OUT_F32, final_act=ACT_NONE, final_out=None, cin=None):
        """nn.Sequential of Conv2d / GELU / (Conv2d + PixelShuffle) holders (stf.py:474-546)."""
        convs = []
        for m in seq:
            if isinstance(m, torch.nn.Conv2d):
                convs.append((m, 0))
            elif isinstance(m, torch.nn.Sequential):  # subpel_conv3x3
                convs.append((m[0], m[1].upscale_factor))
        for k, (m, ps) in enumerate(convs):
            last = k == len(convs) - 1
            pk = self.packed(m, ps)
            if last:
                x = self.conv(x, B, H, W, pk, out=final_out, act=final_act, out_dtype=final_dtype, cin=cin if k == 0 else None)
            else:
                x = self.conv(x, B, H, W, pk, act=ACT_GELU, cin=cin if k == 0 else None, pitch64=True)
            H = (H + 2 * pk.pad - pk.KH) // pk.stride + 1
            W = (W + 2 * pk.pad - pk.KW) // pk.stride + 1
            if ps:
                H, W = H * ps, W * ps
        return x, H, W

    # ---------------------------------------------------------------------------------- WACNN pieces
    def eltwise(self, mode, x, a=None, s=None):
        """mode 0: x*x   1: a*s + x (bf16)   2: a*s + x (fp32)."""
        rows, Cc = x.shape
        out = torch.empty((rows, Cc), dtype=torch.float32 if mode == 2 else torch.bfloat16, device=x.device)
        pa = a.shape[1] if a is not None else 0
        ps = s.shape[1] if s is not None else 0
        check(lib().icm_eltwise_bf16(mode, a.data_ptr() if a is not None else None, pa, s.data_ptr() if s is not None else None, ps,
                                     x.data_ptr(), Cc, out.data_ptr(), Cc, rows, Cc, stream_ptr()), "icm_eltwise_bf16")
        return out

    def gdn(self, x, B, H, W, mod):
        """x bf16 [B*H*W, C] -> x * rsqrt(beta' + gamma' . x^2)  (sqrt for the inverse)  (gdn.py:62-75)."""
        sq = self.eltwise(0, x)
        return self.conv(sq, B, H, W, self.packed(mod), act=ACT_SQRT if mod.inverse else ACT_RSQRT, residual=x, res_mode=RES_MUL)

    def residual_unit(self, x, B, H, W, ru):
        c = ru.conv
        t = self.conv(x, B, H, W, self.packed(c[0]), act=ACT_GELU)
        t = self.conv(t, B, H, W, self.packed(c[2]), act=ACT_GELU)
        return self.conv(t, B, H, W, self.packed(c[4]), act=ACT_GELU, residual=x, res_mode=RES_ADD_BEFORE_ACT)

    def gated_window_block(self, x, B, H, W, blk, out_f32=False):
        """Win_noShift_Attention (layers.py:83-89): x + conv_a(x) * sigmoid(conv_b(x)); x bf16 [B*H*W, C]."""
        Cc = x.shape[1]
        a = x
        for ru in blk.conv_a:
            a = self.residual_unit(a, B, H, W, ru)
        wba = blk.conv_b[0]
        qkv = self.conv(x, B, H, W, self.packed(wba.attn.qkv))
        ao = torch.empty((B * H * W, Cc), dtype=torch.bfloat16, device=x.device)
        check(lib().icm_window_attention_wacnn(qkv.data_ptr(), ao.data_ptr(), self.f32(wba.attn.relative_position_bias_table).data_ptr(),
                                               B, H, W, Cc, wba.num_heads, wba.window_size, wba.shift_size, stream_ptr()),
              "icm_window_attention_wacnn")
        b = self.conv(ao, B, H, W, self.packed(wba.attn.proj), residual=x, res_mode=RES_ADD)
        for ru in list(blk.conv_b)[1:4]:
            b = self.residual_unit(b, B, H, W, ru)
        s = self.conv(b, B, H, W, self.packed(blk.conv_b[4]), act=ACT_SIGMOID)
        return self.eltwise(2 if out_f32 else 1, x, a=a, s=s)
