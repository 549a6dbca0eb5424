"""WACNN (compressai.models.WACNN on the CUDA kernels): op-level checks against fp32 PyTorch / the pinned functional
oracle (oracle/cnn_ref.py), and end-to-end checks against the fixture recorded from the reference model
(tests/golden/cnn_small.npz, 1x3x256x256 = BASELINE.json configs[0]).

Tolerances: operands are bf16, accumulation fp32 -> op-level relative error ~1e-2; x_hat within 0.01 dB PSNR;
strings bit-exact where the inputs are identical (self-consistency, CPU-oracle cross decode)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
PSNR_TOL_DB = 0.01


def psnr(a, b):
    return float(-10 * torch.log10(torch.mean((a.float().cpu() - b.float().cpu()) ** 2)))


def _bf(t):
    return t.bfloat16().float()


def _nhwc(t):
    """[B,C,H,W] fp32 -> bf16 cuda [B*H*W, C]."""
    B, C, H, W = t.shape
    return t.permute(0, 2, 3, 1).reshape(-1, C).contiguous().bfloat16().cuda()


def _nchw(t, B, H, W):
    return t.float().cpu().reshape(B, H, W, -1).permute(0, 3, 1, 2)


def _rel(got, ref):
    return float((got.float().cpu() - ref).norm() / ref.norm())


@pytest.fixture()
def eng():
    from compressai.models._engine import Engine

    return Engine(None)


@pytest.fixture(scope="module")
def sd():
    from oracle import cnn_ref, weights

    return weights.seeded_state_dict(cnn_ref.template_state_dict(), seed=0, stress=True)


@pytest.fixture(scope="module")
def model(sd):
    from compressai.zoo import models

    m = models["cnn"]()
    missing = m.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys
    m.update(force=True)
    return m.cuda().eval()


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "cnn_small.npz"))


@pytest.fixture(scope="module")
def x():
    from oracle import weights

    return weights.seeded_image((1, 3, 256, 256), seed=0)


# ------------------------------------------------------------------------------------------- op level
@pytest.mark.parametrize("inverse", [False, True])
def test_gdn(eng, model, sd, inverse):
    from oracle import cnn_ref

    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 192, 12, 20, generator=g) * 2
    name = "g_s.2" if inverse else "g_a.1"
    mod = model.g_s[2] if inverse else model.g_a[1]
    ref = cnn_ref.gdn(_bf(x), sd, name, inverse)
    got = _nchw(eng.gdn(_nhwc(x), 2, 12, 20, mod), 2, 12, 20)
    assert _rel(got, ref) < 1e-2


@pytest.mark.parametrize("cin,cout", [(320, 192), (192, 192), (192, 3)])
def test_deconv_as_phase_conv(eng, cin, cout):
    from compressai.models._engine import PackedDeconv

    g = torch.Generator().manual_seed(cin + cout)
    x = torch.randn(2, cin, 6, 10, generator=g)
    w = torch.randn(cin, cout, 5, 5, generator=g) / (cin * 6.25) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    ref = F.conv_transpose2d(_bf(x), _bf(w), b, stride=2, padding=2, output_padding=1)
    pk = PackedDeconv(w.cuda(), b.cuda())
    out = eng.conv(_nhwc(x), 2, 6, 10, pk)
    got = _nchw(out, 2, 12, 20)[:, :cout]
    assert got.shape == ref.shape
    err = (got - ref).abs().max()
    assert err < 2e-2 * max(1.0, float(ref.abs().max())), float(err)
    if out.shape[1] > cout:  # padded output channels are written as zeros
        assert float(out[:, cout:].abs().max()) == 0.0


def test_strided_5x5_conv_from_3_channel_image(eng):
    from compressai._native import check, lib, stream_ptr
    from compressai.models._engine import PackedConv

    g = torch.Generator().manual_seed(5)
    x = torch.rand(2, 3, 64, 128, generator=g)
    w = torch.randn(192, 3, 5, 5, generator=g) / 75 ** 0.5
    b = torch.randn(192, generator=g) * 0.1
    ref = F.conv2d(_bf(x), _bf(w), b, stride=2, padding=2)
    xc = x.cuda()
    t = torch.empty((2 * 64 * 128, 8), dtype=torch.bfloat16, device="cuda")
    check(lib().icm_image_to_nhwc(xc.data_ptr(), t.data_ptr(), 2, 3, 64, 128, 8, stream_ptr()))
    assert torch.equal(t[:, :3].float().cpu(), _bf(x).permute(0, 2, 3, 1).reshape(-1, 3)) and float(t[:, 3:].abs().max()) == 0.0
    got = _nchw(eng.conv(t, 2, 64, 128, PackedConv(w.cuda(), b.cuda(), 2, 2, 0)), 2, 32, 64)
    assert (got - ref).abs().max() < 2e-2
    back = torch.empty((2, 3, 64, 128), dtype=torch.float32, device="cuda")
    check(lib().icm_nhwc_to_image(t.data_ptr(), back.data_ptr(), 2, 3, 64, 128, 8, 1, stream_ptr()))
    assert torch.equal(back.cpu(), _bf(x).clamp(0, 1))


@pytest.mark.parametrize("C,win,shift,H,W", [(192, 8, 4, 16, 24), (320, 4, 2, 8, 12), (192, 8, 0, 8, 8)])
def test_shifted_window_attention(eng, C, win, shift, H, W):
    """qkv GEMM -> icm_window_attention_wacnn -> proj GEMM + shortcut  vs  layers/win_attention.py semantics."""
    from compressai.layers import WinBasedAttention
    from oracle import cnn_ref

    g = torch.Generator().manual_seed(C + win)
    blk = WinBasedAttention(dim=C, num_heads=8, window_size=win, shift_size=shift)
    p = {}
    for k, v in blk.state_dict().items():
        if v.is_floating_point():
            v = _bf(torch.randn(v.shape, generator=g) * (0.5 if "table" in k else (1.0 / C ** 0.5 if v.dim() == 2 else 0.1)))
        p[k] = v
    blk.load_state_dict(p)
    blk = blk.cuda()
    x = torch.randn(2, C, H, W, generator=g)
    ref = cnn_ref.shifted_window_attention(_bf(x), p, 8, win, shift)
    from compressai._native import check, lib, stream_ptr

    xt = _nhwc(x)
    qkv = eng.conv(xt, 2, H, W, eng.packed(blk.attn.qkv))
    ao = torch.empty((2 * H * W, C), dtype=torch.bfloat16, device="cuda")
    check(lib().icm_window_attention_wacnn(qkv.data_ptr(), ao.data_ptr(), eng.f32(blk.attn.relative_position_bias_table).data_ptr(),
                                           2, H, W, C, 8, win, shift, stream_ptr()))
    out = eng.conv(ao, 2, H, W, eng.packed(blk.attn.proj), residual=xt, res_mode=0)
    assert _rel(_nchw(out, 2, H, W), ref) < 1.5e-2


def test_residual_unit_and_gated_block(eng, model, sd):
    from oracle import cnn_ref
    from oracle.stf_ref import _sub

    g = torch.Generator().manual_seed(11)
    x = torch.randn(1, 320, 8, 12, generator=g)
    xt = _nhwc(x)
    ru_ref = cnn_ref.residual_unit(_bf(x), _sub(sd, "g_a.8.conv_a.0."))
    ru = eng.residual_unit(xt, 1, 8, 12, model.g_a[8].conv_a[0])
    assert _rel(_nchw(ru, 1, 8, 12), ru_ref) < 1.5e-2
    ref = cnn_ref.gated_window_block(_bf(x), sd, "g_a.8", 8, 4, 2)
    for f32 in (False, True):
        got = eng.gated_window_block(xt, 1, 8, 12, model.g_a[8], out_f32=f32)
        assert got.dtype == (torch.float32 if f32 else torch.bfloat16)
        assert _rel(_nchw(got, 1, 8, 12), ref) < 2e-2


# ------------------------------------------------------------------------------------------- model level
def test_tables_match_reference_update(model, gold):
    eb = model.entropy_bottleneck
    assert np.array_equal(eb.quantized_cdf.cpu().numpy(), gold["eb_cdf"])
    assert np.array_equal(eb.cdf_length.cpu().numpy(), gold["eb_len"]) and np.array_equal(eb.offset.cpu().numpy(), gold["eb_off"])


def test_analysis_and_hyper_transforms_close_to_reference(model, gold, x):
    y, h, w = model._analysis(x.cuda())
    assert (h, w) == (16, 16) and y.dtype == torch.float32
    y_ref = torch.from_numpy(gold["y"]).permute(0, 2, 3, 1).reshape(-1, 320)
    assert _rel(y, y_ref) < 3e-2, _rel(y, y_ref)
    z, zh, zw = model._hyper_analysis(y, 1, h, w)
    z_ref = torch.from_numpy(gold["z"]).permute(0, 2, 3, 1).reshape(-1, 192)
    assert (zh, zw) == (4, 4) and _rel(z, z_ref) < 4e-2, _rel(z, z_ref)


def test_forward_matches_reference_quality_and_rate(model, gold, x):
    out = model(x.cuda())
    assert out["x_hat"].shape == (1, 3, 256, 256) and out["likelihoods"]["y"].shape == (1, 320, 16, 16)
    assert out["likelihoods"]["z"].shape == (1, 192, 4, 4)
    x_ref = torch.from_numpy(gold["x_hat"])
    assert abs(psnr(x, out["x_hat"]) - psnr(x, x_ref)) < PSNR_TOL_DB, (psnr(x, out["x_hat"]), psnr(x, x_ref))
    bits = lambda l: float(-torch.log2(torch.as_tensor(l).float().cpu()).sum())
    by, by_ref = bits(out["likelihoods"]["y"]), bits(gold["y_lik"])
    bz, bz_ref = bits(out["likelihoods"]["z"]), bits(gold["z_lik"])
    assert abs(by - by_ref) / by_ref < 2e-2, (by, by_ref)
    assert abs(bz - bz_ref) / bz_ref < 2e-2, (bz, bz_ref)


def test_compress_decompress_self_consistency(model, gold, x):
    xc = x.cuda()
    c = model.compress(xc)
    assert list(c["shape"]) == [4, 4] and len(c["strings"][0]) == 1 and len(c["strings"][1]) == 1
    d = model.decompress(c["strings"], c["shape"])
    f = model(xc)
    assert torch.equal(d["x_hat"], f["x_hat"].clamp(0, 1))
    ny, nz = len(c["strings"][0][0]), len(c["strings"][1][0])
    assert abs(ny - gold["y_string"].size) / gold["y_string"].size < 3e-2, (ny, gold["y_string"].size)
    assert abs(nz - gold["z_string"].size) / gold["z_string"].size < 5e-2, (nz, gold["z_string"].size)
    assert abs(psnr(x, d["x_hat"]) - psnr(x, torch.from_numpy(gold["x_hat"]).clamp(0, 1))) < PSNR_TOL_DB


def test_batch_strings_equal_single_image_strings(model):
    from oracle import weights

    xs = torch.cat([weights.seeded_image((1, 3, 64, 128), seed=s) for s in (1, 2, 3)]).cuda()
    cb = model.compress(xs)
    for b in range(3):
        c1 = model.compress(xs[b : b + 1])
        assert c1["strings"][0][0] == cb["strings"][0][b] and c1["strings"][1][0] == cb["strings"][1][b]
    d = model.decompress(cb["strings"], cb["shape"])
    assert torch.equal(d["x_hat"], model(xs)["x_hat"].clamp(0, 1))


def test_strings_decode_with_the_cpu_oracle(model, x):
    """The GPU-produced y-string decodes, with the pinned CPU coder and the GPU-side indexes, to the symbols the
    GPU encoder consumed, and the CPU coder produces the same bytes from them."""
    from compressai._native import NULL_VIEW, check, lib, stream_ptr, view_bcp
    from oracle import coder, entropy

    xc = x.cuda()
    B = 1
    y, h, w = model._analysis(xc)
    z, zh, zw = model._hyper_analysis(y, B, h, w)
    c = model.compress(xc)
    eb = model.entropy_bottleneck
    Pz = zh * zw
    z_sym = torch.empty((B, 192 * Pz), dtype=torch.int32, device="cuda")
    z_idx = torch.empty_like(z_sym)
    z_hat = torch.empty((B * Pz, 192), dtype=torch.bfloat16, device="cuda")
    check(lib().icm_eb_process(0, view_bcp(z, B, 192, Pz), B, 192, Pz, eb.packed_params().data_ptr(), 0.0, z_sym.data_ptr(),
                               z_idx.data_ptr(), NULL_VIEW, view_bcp(z_hat, B, 192, Pz), NULL_VIEW, stream_ptr()))
    ms, ss = model._hyper_synthesis(z_hat, B, zh, zw)
    _, sym, idx = model._slice_loop("compress", B, h, w, ms, ss, y=y)
    cdf, lengths, offsets = entropy.gc_tables()
    got = coder.RansDecoder().decode_with_indexes(c["strings"][0][0], idx[0].cpu().numpy(), cdf, lengths, offsets)
    assert np.array_equal(got, sym[0].cpu().numpy())
    assert coder.rans_encode(sym[0].cpu().numpy(), idx[0].cpu().numpy(), cdf, lengths, offsets) == c["strings"][0][0]
    assert int(idx.unique().numel()) > 20 and int(sym.abs().max()) > 8


def test_input_validation(model):
    with pytest.raises(ValueError):
        model.compress(torch.zeros(1, 3, 100, 128, device="cuda"))


def test_config4_wacnn2_codec_at_832x1216(sd, golden_dir):
    """BASELINE.json configs[3] as SURVEY.md §8d reads it: the WACNN2 codec on a 800x1216 image zero-padded to
    832x1216 (y 320x52x76 = 1 264 640 symbols, z 192x13x19).  Too large for the CPU oracle inside a test, so the
    size-independent properties are checked: decompress(compress(x)) == clamp(forward(x).x_hat) exactly, the strings of
    the padded image decode on the pinned CPU coder to the symbols the GPU encoded, and cropping is the caller's job."""
    import torch.nn.functional as F

    from compressai.zoo import models
    from oracle import coder, entropy, weights

    m = models["cnn2"]()
    assert not m.load_state_dict(sd, strict=False).unexpected_keys
    m.update(force=True)
    m = m.cuda().eval()
    x = weights.seeded_image((1, 3, 800, 1216), seed=5)
    xp = F.pad(x, (0, 0, 16, 16)).cuda()  # eval_model/__main__.py:103-115 pads to a multiple of 64, centred
    assert xp.shape == (1, 3, 832, 1216)
    c = m.compress(xp)
    assert list(c["shape"]) == [13, 19]
    d = m.decompress(c["strings"], c["shape"])
    f = m(xp)
    assert f["likelihoods"]["y"].shape == (1, 320, 52, 76) and f["likelihoods"]["z"].shape == (1, 192, 13, 19)
    assert torch.equal(d["x_hat"], f["x_hat"].clamp(0, 1))
    # stream sizes against the reference's own run of this image (digests in tests/golden/cnn2_full.json; the transforms run
    # with bf16 operands here, so the bytes differ and only the sizes are comparable)
    import json

    gold = json.load(open(os.path.join(golden_dir, "cnn2_full.json")))
    ny, nz = len(c["strings"][0][0]), len(c["strings"][1][0])
    assert abs(ny - gold["y_bytes"]) / gold["y_bytes"] < 3e-2, (ny, gold["y_bytes"])
    assert abs(nz - gold["z_bytes"]) / gold["z_bytes"] < 5e-2, (nz, gold["z_bytes"])
    # the y-string is a valid reference-format stream of exactly 1 264 640 symbols: decode it on the CPU with the
    # indexes of the GPU slice loop and re-encode
    y, h, w = m._analysis(xp)
    z, zh, zw = m._hyper_analysis(y, 1, h, w)
    from compressai._native import NULL_VIEW, check, lib, stream_ptr, view_bcp

    Pz = zh * zw
    z_sym = torch.empty((1, 192 * Pz), dtype=torch.int32, device="cuda")
    z_idx = torch.empty_like(z_sym)
    z_hat = torch.empty((Pz, 192), dtype=torch.bfloat16, device="cuda")
    check(lib().icm_eb_process(0, view_bcp(z, 1, 192, Pz), 1, 192, Pz, m.entropy_bottleneck.packed_params().data_ptr(), 0.0, z_sym.data_ptr(),
                               z_idx.data_ptr(), NULL_VIEW, view_bcp(z_hat, 1, 192, Pz), NULL_VIEW, stream_ptr()))
    ms, ss = m._hyper_synthesis(z_hat, 1, zh, zw)
    _, sym, idx = m._slice_loop("compress", 1, h, w, ms, ss, y=y)
    assert sym.shape[1] == 1264640
    cdf, lengths, offsets = entropy.gc_tables()
    got = coder.RansDecoder().decode_with_indexes(c["strings"][0][0], idx[0].cpu().numpy(), cdf, lengths, offsets)
    assert np.array_equal(got, sym[0].cpu().numpy())
    assert coder.rans_encode(sym[0].cpu().numpy(), idx[0].cpu().numpy(), cdf, lengths, offsets) == c["strings"][0][0]
