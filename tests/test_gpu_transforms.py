"""Transform kernels (csrc/conv.cu tcgen05 implicit GEMM, csrc/transforms.cu) vs a plain fp32 PyTorch
reference of the same op on the CPU.  Operands are bf16 (rounded identically on both sides), accumulation
fp32, so the tolerance covers accumulation order and the bf16 rounding of the stored output only."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture()
def eng():
    from compressai.models._engine import Engine

    return Engine(None)


def _bf(t):
    return t.bfloat16().float()


def _close(got, ref, rtol, atol):
    got, ref = got.float().cpu(), ref.float().cpu()
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    assert torch.all(err <= tol), f"max err {float(err.max()):.4g} at ref scale {float(ref.abs().max()):.4g}; worst excess {float((err - tol).max()):.4g}"


CONV_CASES = [
    # B, H, W, Cin, Cout, k, stride, pad, act, out_dtype, pixel_shuffle, pitch_extra
    (1, 8, 12, 384, 224, 3, 1, 1, 1, 0, 0, 0),      # first conv of a cc stack at 128x192
    (2, 8, 12, 416, 224, 3, 1, 1, 1, 0, 0, 192),    # Cin not a multiple of 64, wider pitch (support buffer)
    (1, 48, 32, 64, 32, 3, 1, 1, 2, 1, 0, 0),       # last conv of an lrp stack: 0.5*tanh, fp32 out
    (2, 12, 8, 336, 288, 3, 2, 1, 1, 0, 0, 0),      # h_a stride-2 conv
    (1, 3, 5, 192, 240, 3, 1, 1, 1, 0, 0, 0),       # tiny spatial size (z plane)
    (1, 6, 4, 240, 1152, 3, 1, 1, 1, 0, 2, 0),      # subpel conv + PixelShuffle(2)
    (1, 20, 24, 48, 192, 5, 1, 2, 0, 0, 2, 0),      # end_conv.0 5x5 + PixelShuffle(2)
    (1, 1, 1000, 48, 144, 1, 1, 0, 0, 0, 0, 0),     # qkv linear, C=48 (K < 64)
    (1, 1, 777, 768, 384, 1, 1, 0, 0, 1, 0, 0),     # PatchMerging reduction, ragged M
    (1, 1, 300, 1536, 384, 1, 1, 0, 0, 1, 0, 0),    # fc2 at C=384
]


@pytest.mark.parametrize("case", CONV_CASES, ids=[f"c{i}" for i in range(len(CONV_CASES))])
def test_conv_igemm(eng, case):
    from compressai.models._engine import PackedConv

    B, H, W, Cin, Cout, k, stride, pad, act, odt, ps, extra = case
    g = torch.Generator().manual_seed(hash(case) % 1000)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    ref = F.conv2d(_bf(x), _bf(w), b, stride=stride, padding=pad)
    if act == 1:
        ref = F.gelu(ref)
    elif act == 2:
        ref = 0.5 * torch.tanh(ref)
    if ps:
        ref = F.pixel_shuffle(ref, ps)
    pitch = Cin + extra
    xin = torch.full((B * H * W, pitch), 7.0)  # channels beyond Cin must be ignored
    xin[:, :Cin] = x.permute(0, 2, 3, 1).reshape(-1, Cin)
    pk = PackedConv(w.cuda(), b.cuda(), stride, pad, ps)
    out = eng.conv(xin.cuda().bfloat16(), B, H, W, pk, act=act, out_dtype=odt)
    torch.cuda.synchronize()
    ref_cl = ref.permute(0, 2, 3, 1).reshape(-1, ref.shape[1])
    assert out.shape == ref_cl.shape
    _close(out, ref_cl, rtol=1e-2 if odt == 0 else 2e-3, atol=2e-3)


def test_conv_residual_and_offset_output(eng):
    from compressai.models._engine import PackedConv

    g = torch.Generator().manual_seed(1)
    M, K, N = 500, 192, 96
    x, w, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
    res = torch.randn(M, N, generator=g)
    pk = PackedConv(w.cuda(), b.cuda())
    r = res.cuda().clone()
    eng.linear(x.cuda().bfloat16(), pk, out=r, out_dtype=1, residual=r)  # in place, like the Swin residual stream
    _close(r, F.linear(_bf(x), _bf(w), b) + res, rtol=2e-3, atol=2e-3)
    wide = torch.zeros(M, 256, dtype=torch.bfloat16, device="cuda")
    eng.linear(x.cuda().bfloat16(), pk, out=wide, out_offset=64)
    assert torch.all(wide[:, :64] == 0) and torch.all(wide[:, 160:] == 0)
    _close(wide[:, 64:160], F.linear(_bf(x), _bf(w), b), rtol=1e-2, atol=2e-3)


def test_conv_is_batch_invariant(eng):
    """Encoder/decoder determinism: the same pixel gives bit-identical results alone or inside a batch."""
    from compressai.models._engine import PackedConv

    g = torch.Generator().manual_seed(2)
    x = torch.randn(4 * 8 * 12, 448, generator=g).cuda().bfloat16()
    pk = PackedConv((torch.randn(224, 448, 3, 3, generator=g) / 60).cuda(), torch.randn(224, generator=g).cuda(), 1, 1)
    full = eng.conv(x, 4, 8, 12, pk, out_dtype=1)
    one = eng.conv(x[2 * 96:3 * 96].contiguous(), 1, 8, 12, pk, out_dtype=1)
    assert torch.equal(full[2 * 96:3 * 96], one)


def _packed_group(pks):
    from compressai.models._engine import PackedGroup

    return PackedGroup(pks)


def test_grouped_conv_is_bit_identical_to_separate_launches(eng):
    """icm_conv2d_grouped: (a) G stacks reading ONE shared input, each with its own tail chunk ([support | y_hat_i] of the LRP
    stacks of slices >= max_support, stf.py:623-626), (b) stacked inputs (layers 2..5 of grouped stacks, cc_mean || cc_scale),
    (c) PixelShuffle outputs (h_mean_s || h_scale_s) -- every output element must equal the ungrouped kernel's bit for bit,
    because encoder and decoder may mix the two forms."""
    from compressai.models._engine import PackedConv

    g = torch.Generator().manual_seed(11)
    B, H, W = 3, 8, 12
    # (a) Cin = 128 support channels + 32 tail channels at per-group positions of a 320-wide row
    G, pitch, Cs = 3, 320, 128
    x = torch.randn(B * H * W, pitch, generator=g).cuda().bfloat16()
    pks = [PackedConv((torch.randn(80, Cs + 32, 3, 3, generator=g) / 40).cuda(), torch.randn(80, generator=g).cuda(), 1, 1) for _ in range(G)]
    tails = [Cs + 32 * (k + 1) for k in range(G)]
    got = eng.conv_group(x, B, H, W, _packed_group(pks), B, [0] * G, tail_channels=tails, act=1, out_dtype=1)
    for k in range(G):
        xin = torch.cat([x[:, :Cs], x[:, tails[k]:tails[k] + 32]], 1).contiguous()
        assert torch.equal(got[k], eng.conv(xin, B, H, W, pks[k], act=1, out_dtype=1)), f"shared input + tail, group {k}"
    # (b) stacked inputs with image offsets, outputs interleaved into one wide row (mu / scale of several slices side by side)
    xs = torch.randn(2 * B * H * W, 96, generator=g).cuda().bfloat16()
    pk2 = [PackedConv((torch.randn(32, 96, 3, 3, generator=g) / 30).cuda(), torch.randn(32, generator=g).cuda(), 1, 1) for _ in range(2)]
    wide = torch.zeros(B * H * W, 64, dtype=torch.float32, device="cuda")
    eng.conv_group(xs, B, H, W, _packed_group(pk2), 2 * B, [0, B], out=wide, out_group_stride=32, act=2, out_dtype=1)
    for k in range(2):
        one = eng.conv(xs[k * B * H * W:(k + 1) * B * H * W].contiguous(), B, H, W, pk2[k], act=2, out_dtype=1)
        assert torch.equal(wide[:, 32 * k:32 * k + 32], one), f"stacked input, group {k}"
    # (c) PixelShuffle(2) outputs, shared input, bf16
    xz = torch.randn(B * 3 * 4, 240, generator=g).cuda().bfloat16()
    pk3 = [PackedConv((torch.randn(1152, 240, 3, 3, generator=g) / 45).cuda(), torch.randn(1152, generator=g).cuda(), 1, 1, 2) for _ in range(2)]
    got = eng.conv_group(xz, B, 3, 4, _packed_group(pk3), B, [0, 0], act=1)
    for k in range(2):
        assert torch.equal(got[k], eng.conv(xz, B, 3, 4, pk3[k], act=1)), f"pixel shuffle, group {k}"


@pytest.mark.parametrize("C,gather", [(48, False), (96, False), (384, False), (192, True), (768, True)])
def test_layernorm(eng, C, gather):
    g = torch.Generator().manual_seed(C)
    norm = torch.nn.LayerNorm(C)
    with torch.no_grad():
        norm.weight.copy_(1 + 0.1 * torch.randn(C, generator=g))
        norm.bias.copy_(0.1 * torch.randn(C, generator=g))
    if gather:
        B, H, W, Cs = 2, 6, 8, C // 4
        x = torch.randn(B, H, W, Cs, generator=g) * 3 + 1
        cat = torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], -1).reshape(-1, C)
        ref = norm(cat)
        out = eng.layernorm(x.reshape(-1, Cs).cuda(), norm.cuda(), out_dtype=1, gather=(B, H, W))
    else:
        x = torch.randn(333, C, generator=g) * 3 + 1
        ref = norm(x)
        out = eng.layernorm(x.cuda(), norm.cuda(), out_dtype=1)
    _close(out, ref.detach(), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("heads,shift", [(3, 0), (3, 2), (6, 2), (24, 2), (12, 0)])
def test_window_attention(eng, heads, shift):
    from oracle import stf_ref

    C = heads * 16
    B, H, W = 2, 8, 12
    g = torch.Generator().manual_seed(heads * 10 + shift)
    qkv = _bf(torch.randn(B, H, W, 3 * C, generator=g))
    table = torch.randn(49, heads, generator=g) * 0.5
    # reference: the reference's own sequence roll -> partition -> softmax(qk^T*scale + bias + mask) v -> reverse -> roll
    x = qkv
    if shift:
        x = torch.roll(x, (-shift, -shift), (1, 2))
    xw = x.reshape(B, H // 4, 4, W // 4, 4, 3 * C).permute(0, 1, 3, 2, 4, 5).reshape(-1, 16, 3, heads, 16).permute(2, 0, 3, 1, 4)
    q, k, v = xw[0] * 0.25, xw[1], xw[2]
    a = q @ k.transpose(-2, -1) + table[stf_ref.rel_pos_index(4).reshape(-1)].reshape(16, 16, heads).permute(2, 0, 1)[None]
    if shift:
        mask = stf_ref.shift_mask(H, W, 4)
        a = (a.reshape(B, -1, heads, 16, 16) + mask[None, :, None]).reshape(-1, heads, 16, 16)
    o = (torch.softmax(a, -1) @ v).transpose(1, 2).reshape(-1, 16, C)
    o = o.reshape(B, H // 4, W // 4, 4, 4, C).permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, C)
    if shift:
        o = torch.roll(o, (shift, shift), (1, 2))
    got = eng.window_attention(qkv.reshape(-1, 3 * C).cuda().bfloat16(), B, H, W, C, heads, 4, shift, table.cuda())
    _close(got, o.reshape(-1, C), rtol=1e-2, atol=5e-3)


def test_patch_embed_and_final_conv(eng):
    from compressai._native import check, lib, stream_ptr

    g = torch.Generator().manual_seed(4)
    B, H, W = 2, 44, 70  # not multiples of the 32x32 CTA tile of the output convolution
    img = torch.rand(B, 3, H, W, generator=g)
    w, b = torch.randn(48, 3, 2, 2, generator=g) * 0.3, torch.randn(48, generator=g) * 0.1
    gam, bet = 1 + 0.1 * torch.randn(48, generator=g), 0.1 * torch.randn(48, generator=g)
    t = F.conv2d(img, w, b, stride=2).flatten(2).transpose(1, 2)
    ref = F.layer_norm(t, (48,), gam, bet, 1e-5).reshape(-1, 48)
    out = torch.empty(B * (H // 2) * (W // 2), 48, device="cuda")
    c = lambda v: v.cuda().contiguous()
    ci, cw, cb, cg, cbe = c(img), c(w), c(b), c(gam), c(bet)
    check(lib().icm_patch_embed(ci.data_ptr(), cw.data_ptr(), cb.data_ptr(), cg.data_ptr(), cbe.data_ptr(), out.data_ptr(), B, H, W, 48, stream_ptr()))
    _close(out, ref, rtol=1e-4, atol=1e-4)
    x = torch.randn(B, 48, H, W, generator=g)
    w3, b3 = torch.randn(3, 48, 3, 3, generator=g) * 0.1, torch.randn(3, generator=g) * 0.1
    ref = F.conv2d(_bf(x), _bf(w3), b3, padding=1)  # bf16 operands (tensor-core kernel), fp32 accumulation
    xin = c(x.permute(0, 2, 3, 1).reshape(-1, 48)).bfloat16()
    for clamp in (0, 1):
        o = torch.empty(B, 3, H, W, device="cuda")
        cw3, cb3 = c(w3), c(b3)
        check(lib().icm_final_conv(xin.data_ptr(), cw3.data_ptr(), cb3.data_ptr(), o.data_ptr(), B, H, W, 48, clamp, stream_ptr()))
        _close(o, ref.clamp(0, 1) if clamp else ref, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("C,M", [(48, 1000), (96, 777), (192, 4096 + 57), (48, 128 * 300 + 5)])
def test_fused_swin_mlp_is_bit_identical_to_two_launches(eng, C, M):
    """icm_swin_mlp (hidden activation kept on chip) == fc1 + GELU launch followed by fc2 + residual launch, bit for bit
    (same K order, same rounding points), and close to the fp32 PyTorch MLP."""
    from compressai.models._engine import PackedConv  # noqa: F401

    g = torch.Generator().manual_seed(C + M)
    mlp = torch.nn.Module()
    mlp.fc1 = torch.nn.Linear(C, 4 * C)
    mlp.fc2 = torch.nn.Linear(4 * C, C)
    with torch.no_grad():
        for prm in mlp.parameters():
            prm.copy_(torch.randn(prm.shape, generator=g) * (0.1 if prm.dim() == 1 else 1.0 / prm.shape[1] ** 0.5))
    mlp = mlp.cuda()
    x0 = torch.randn(M, C, generator=g) * 2
    xn = (torch.randn(M, C, generator=g) * 1.5).cuda().bfloat16()
    xa, xb = x0.cuda().clone(), x0.cuda().clone()
    eng.fused_mlp = True
    eng.mlp(xn, xa, mlp)
    eng.fused_mlp = False
    eng.mlp(xn, xb, mlp)
    torch.cuda.synchronize()
    assert torch.equal(xa, xb), float((xa - xb).abs().max())
    with torch.no_grad():
        h = F.gelu(F.linear(xn.float().cpu(), _bf(mlp.fc1.weight.cpu()), mlp.fc1.bias.cpu()))
        ref = x0 + F.linear(_bf(h), _bf(mlp.fc2.weight.cpu()), mlp.fc2.bias.cpu())
    _close(xa, ref, rtol=5e-3, atol=5e-3)


@pytest.mark.parametrize("C,heads,shifted,B,H,W", [(48, 3, False, 2, 8, 12), (48, 3, True, 2, 8, 12), (96, 6, True, 1, 12, 8), (96, 6, False, 3, 4, 4),
                                                   (48, 3, True, 1, 64, 32)])
def test_fused_swin_block_against_the_fp32_block_and_the_six_launch_path(eng, C, heads, shifted, B, H, W):
    """icm_swin_block (csrc/swin_fused.cu: LayerNorm -> qkv -> window attention -> proj -> +x -> LayerNorm -> MLP -> +x in
    registers, one warp per window) against the oracle's fp32 restatement of SwinTransformerBlock.forward (stf.py:149-199)
    and against the six-launch path it replaces (same bf16 rounding points; only the accumulation order differs)."""
    from oracle import stf_ref

    g = torch.Generator().manual_seed(C * 7 + heads + int(shifted))
    blk = torch.nn.Module()
    blk.window_size = 4
    blk.norm1, blk.norm2 = torch.nn.LayerNorm(C), torch.nn.LayerNorm(C)
    blk.attn = torch.nn.Module()
    blk.attn.num_heads = heads
    blk.attn.qkv, blk.attn.proj = torch.nn.Linear(C, 3 * C), torch.nn.Linear(C, C)
    blk.attn.relative_position_bias_table = torch.nn.Parameter(torch.randn(49, heads, generator=g) * 0.5)
    blk.mlp = torch.nn.Module()
    blk.mlp.fc1, blk.mlp.fc2 = torch.nn.Linear(C, 4 * C), torch.nn.Linear(4 * C, C)
    with torch.no_grad():
        for name, prm in blk.named_parameters():
            if "relative_position" in name:
                continue
            if "norm" in name:
                prm.copy_((1.0 if name.endswith("weight") else 0.0) + 0.1 * torch.randn(prm.shape, generator=g))
            else:
                prm.copy_(torch.randn(prm.shape, generator=g) * (0.1 if prm.dim() == 1 else 1.0 / prm.shape[1] ** 0.5))
    x0 = torch.randn(B * H * W, C, generator=g) * 2 + 0.3
    # fp32 reference: the oracle's swin_block on [B, L, C] tokens
    p = {k: v.detach() for k, v in blk.state_dict().items()}
    p["attn.relative_position_index"] = stf_ref.rel_pos_index(4)
    mask = stf_ref.shift_mask(H, W, 4) if shifted else None
    ref = stf_ref.swin_block(x0.reshape(B, H * W, C), H, W, p, heads, shifted, mask).reshape(-1, C)
    blk = blk.cuda()
    xa, xb = x0.cuda().clone(), x0.cuda().clone()
    eng.fused_block = True
    eng.swin_block(xa, B, H, W, blk, shifted)
    eng.fused_block = False
    eng.swin_block(xb, B, H, W, blk, shifted)
    eng.fused_block = True
    torch.cuda.synchronize()
    _close(xa, ref, rtol=2e-2, atol=2e-2)   # bf16 operands
    _close(xa, xb, rtol=1e-2, atol=1e-2)    # same rounding points as the unfused path; the MLP's GELU is the tanh form here and
                                            # the erf-polynomial form there (|difference| <= 5e-4 before the bf16 rounding)
    # the halves on their own (C = 96 always runs them as two launches)
    from compressai._native import check, lib, stream_ptr

    xc = x0.cuda().clone()
    at, mlp = blk.attn, blk.mlp
    qkv, proj, f1, f2 = eng.packed(at.qkv), eng.packed(at.proj), eng.packed(mlp.fc1), eng.packed(mlp.fc2)
    args = (qkv.w.data_ptr(), qkv.bias.data_ptr(), proj.w.data_ptr(), proj.bias.data_ptr(), eng.f32(at.relative_position_bias_table).data_ptr(),
            eng.f32(blk.norm1.weight).data_ptr(), eng.f32(blk.norm1.bias).data_ptr(), f1.w.data_ptr(), f1.bias.data_ptr(), f2.w.data_ptr(),
            f2.bias.data_ptr(), eng.f32(blk.norm2.weight).data_ptr(), eng.f32(blk.norm2.bias).data_ptr(), stream_ptr())
    for parts in (1, 2):
        check(lib().icm_swin_block(xc.data_ptr(), B, H, W, C, heads, 4, 2 if shifted else 0, parts, *args))
    torch.cuda.synchronize()
    assert torch.equal(xc, xa)
