"""Functional fp32 CPU restatement of the reference STF transforms + slice loop (rows T1-T10).

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Works straight from a reference-named state_dict
(SURVEY.md Appendix B); pinned against the real reference modules by oracle/make_golden.py
(tests/golden/stf_small.npz) and tests/test_oracle_pinned.py.  All line numbers refer to
/root/reference/compressai/models/stf.py unless noted.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import coder, entropy

WINDOW = 4
DEPTHS = (2, 2, 6, 2)
HEADS = (3, 6, 12, 24)
NUM_SLICES = 12
MAX_SUPPORT = 6


def _sub(sd, prefix):
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


def layer_norm(x, p, name):
    return F.layer_norm(x, (x.shape[-1],), p[name + ".weight"], p[name + ".bias"], 1e-5)


def rel_pos_index(w):
    """:69-80."""
    ii = torch.arange(w * w)
    ih, iw = ii // w, ii % w
    return (ih[:, None] - ih[None, :] + w - 1) * (2 * w - 1) + (iw[:, None] - iw[None, :] + w - 1)


def shift_mask(Hp, Wp, w):
    """:316-334: region label 3*r+c on the padded grid; 0 where labels agree else -100."""
    s = w // 2
    lab = torch.zeros(Hp, Wp)
    bands_h = ((0, Hp - w), (Hp - w, Hp - s), (Hp - s, Hp))
    bands_w = ((0, Wp - w), (Wp - w, Wp - s), (Wp - s, Wp))
    for r, (h0, h1) in enumerate(bands_h):
        for c, (w0, w1) in enumerate(bands_w):
            lab[h0:h1, w0:w1] = 3 * r + c
    win = lab.reshape(Hp // w, w, Wp // w, w).permute(0, 2, 1, 3).reshape(-1, w * w)
    diff = win[:, None, :] - win[:, :, None]
    return torch.where(diff != 0, torch.tensor(-100.0), torch.tensor(0.0))


def window_attention(xw, p, heads, mask):
    """:90-121 on xw [nW*B, 16, C]."""
    Bw, N, C = xw.shape
    hd = C // heads
    qkv = F.linear(xw, p["attn.qkv.weight"], p["attn.qkv.bias"]).reshape(Bw, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * hd ** -0.5, qkv[1], qkv[2]
    a = q @ k.transpose(-2, -1)
    bias = p["attn.relative_position_bias_table"][rel_pos_index(WINDOW).reshape(-1)].reshape(N, N, heads).permute(2, 0, 1)
    a = a + bias.unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        a = (a.reshape(Bw // nW, nW, heads, N, N) + mask[None, :, None]).reshape(-1, heads, N, N)
    a = torch.softmax(a, -1)
    out = (a @ v).transpose(1, 2).reshape(Bw, N, C)
    return F.linear(out, p["attn.proj.weight"], p["attn.proj.bias"])


def swin_block(x, H, W, p, heads, shifted, mask):
    """:149-199 (drop_path is identity in eval)."""
    B, L, C = x.shape
    w = WINDOW
    h = layer_norm(x, p, "norm1").reshape(B, H, W, C)
    pr, pb = (w - W % w) % w, (w - H % w) % w
    h = F.pad(h, (0, 0, 0, pr, 0, pb))
    Hp, Wp = H + pb, W + pr
    if shifted:
        h = torch.roll(h, (-(w // 2), -(w // 2)), (1, 2))
    xw = h.reshape(B, Hp // w, w, Wp // w, w, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, w * w, C)
    aw = window_attention(xw, p, heads, mask if shifted else None)
    h = aw.reshape(B, Hp // w, Wp // w, w, w, C).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, C)
    if shifted:
        h = torch.roll(h, (w // 2, w // 2), (1, 2))
    h = h[:, :H, :W].reshape(B, L, C)
    x = x + h
    m = layer_norm(x, p, "norm2")
    m = F.linear(F.gelu(F.linear(m, p["mlp.fc1.weight"], p["mlp.fc1.bias"])), p["mlp.fc2.weight"], p["mlp.fc2.bias"])
    return x + m


def patch_merge(x, H, W, p):
    """:209-235."""
    B, L, C = x.shape
    x = x.reshape(B, H, W, C)
    if H % 2 or W % 2:
        x = F.pad(x, (0, 0, 0, W % 2, 0, H % 2))
    x = torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], -1)
    x = x.reshape(B, -1, 4 * C)
    x = layer_norm(x, p, "downsample.norm")
    return F.linear(x, p["downsample.reduction.weight"]), (H + 1) // 2, (W + 1) // 2


def patch_split(x, H, W, p):
    """:251-260."""
    B, L, C = x.shape
    x = F.linear(layer_norm(x, p, "downsample.norm"), p["downsample.reduction.weight"])
    x = F.pixel_shuffle(x.permute(0, 2, 1).reshape(B, 2 * C, H, W), 2)
    return x.permute(0, 2, 3, 1).reshape(B, 4 * L, C // 2), 2 * H, 2 * W


def basic_layer(x, H, W, p, depth, heads, kind):
    """:308-347."""
    w = WINDOW
    Hp, Wp = -(-H // w) * w, -(-W // w) * w
    mask = shift_mask(Hp, Wp, w)
    for i in range(depth):
        x = swin_block(x, H, W, _sub(p, f"blocks.{i}."), heads, i % 2 == 1, mask)
    if kind == "merge":
        return patch_merge(x, H, W, p)
    if kind == "split":
        return patch_split(x, H, W, p)
    return x, H, W


def analysis(sd, x):
    """patch_embed (:365-381) + layers (:584-595): image [B,3,H,W] -> y [B,384,H/16,W/16]."""
    if x.shape[3] % 2:
        x = F.pad(x, (0, 1))
    if x.shape[2] % 2:
        x = F.pad(x, (0, 0, 0, 1))
    t = F.conv2d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=2)
    B, C, H, W = t.shape
    t = t.flatten(2).transpose(1, 2)
    t = F.layer_norm(t, (C,), sd["patch_embed.norm.weight"], sd["patch_embed.norm.bias"], 1e-5)
    for i in range(4):
        t, H, W = basic_layer(t, H, W, _sub(sd, f"layers.{i}."), DEPTHS[i], HEADS[i], "merge" if i < 3 else None)
    return t.reshape(B, H, W, -1).permute(0, 3, 1, 2).contiguous()


def synthesis(sd, y_hat):
    """:636-641: y_hat [B,384,h,w] -> x_hat [B,3,16h,16w] (unclamped)."""
    B, C, H, W = y_hat.shape
    t = y_hat.permute(0, 2, 3, 1).reshape(B, H * W, C)
    for i in range(4):
        t, H, W = basic_layer(t, H, W, _sub(sd, f"syn_layers.{i}."), DEPTHS[3 - i], HEADS[3 - i], "split" if i < 3 else None)
    t = t.reshape(B, H, W, -1).permute(0, 3, 1, 2)
    t = F.conv2d(t, sd["end_conv.0.weight"], sd["end_conv.0.bias"], padding=2)
    t = F.pixel_shuffle(t, 2)
    return F.conv2d(t, sd["end_conv.2.weight"], sd["end_conv.2.bias"], padding=1)


def conv_stack(sd, prefix, x, strides=(1, 1, 1, 1, 1)):
    """Sequential(conv3x3, GELU, ... , conv3x3) with module indices 0,2,4,6,8 (:474-484,508-546)."""
    for k, s in enumerate(strides):
        x = F.conv2d(x, sd[f"{prefix}.{2 * k}.weight"], sd[f"{prefix}.{2 * k}.bias"], stride=s, padding=1)
        if k < len(strides) - 1:
            x = F.gelu(x)
    return x


def hyper_synthesis(sd, prefix, z_hat):
    """h_mean_s / h_scale_s (:486-507): conv, subpel, conv, subpel, conv with GELU between."""
    c = lambda t, n: F.conv2d(t, sd[f"{prefix}.{n}.weight"], sd[f"{prefix}.{n}.bias"], padding=1)
    t = F.gelu(c(z_hat, "0"))
    t = F.gelu(F.pixel_shuffle(c(t, "2.0"), 2))
    t = F.gelu(c(t, "4"))
    t = F.gelu(F.pixel_shuffle(c(t, "6.0"), 2))
    return c(t, "8")


def eb_params(sd):
    return _sub(sd, "entropy_bottleneck.")


def slice_loop(sd, y, latent_means, latent_scales, mode, decoder=None, tables=None, num_slices=NUM_SLICES, max_support=MAX_SUPPORT):
    """:607-631 (forward), :703-726 (compress), :754-776 (decompress).

    mode "forward": returns (y_hat, y_likelihoods); "compress": returns (y_hat, symbols, indexes) with
    symbols/indexes lists per slice [B,32,h,w]; "decompress": y is None, symbols come from `decoder`.
    """
    table = entropy.scale_table()
    y_slices = y.chunk(num_slices, 1) if y is not None else [None] * num_slices
    hats, liks, syms, idxs = [], [], [], []
    for i in range(num_slices):
        support = hats[:max_support]
        mean_support = torch.cat([latent_means] + support, 1)
        mu = conv_stack(sd, f"cc_mean_transforms.{i}", mean_support)
        scale_support = torch.cat([latent_scales] + support, 1)
        scale = conv_stack(sd, f"cc_scale_transforms.{i}", scale_support)
        if mode == "forward":
            _, lik = entropy.gc_forward_eval(y_slices[i], scale, mu)
            liks.append(lik)
            y_hat = torch.round(y_slices[i] - mu) + mu
        elif mode == "compress":
            idx = entropy.build_indexes(scale, table)
            q = entropy.quantize_symbols(y_slices[i], mu)
            y_hat = q + mu
            syms.append(q)
            idxs.append(idx)
        else:
            idx = entropy.build_indexes(scale, table)
            cdf, lengths, offsets = tables
            rv = decoder.decode_stream(idx.reshape(-1).numpy(), cdf, lengths, offsets)
            y_hat = entropy.dequantize(torch.from_numpy(rv.astype(np.float32)).reshape(mu.shape), mu)
            idxs.append(idx)
        lrp = conv_stack(sd, f"lrp_transforms.{i}", torch.cat([mean_support, y_hat], 1))
        hats.append(y_hat + 0.5 * torch.tanh(lrp))
    y_hat = torch.cat(hats, 1)
    if mode == "forward":
        return y_hat, torch.cat(liks, 1)
    return y_hat, syms, idxs


@torch.no_grad()
def forward(sd, x):
    """:582-645 in eval mode."""
    y = analysis(sd, x)
    z = conv_stack(sd, "h_a", y, (1, 1, 2, 1, 2))
    z_hat, z_lik = entropy.eb_forward_eval(eb_params(sd), z)
    ls, lm = hyper_synthesis(sd, "h_scale_s", z_hat), hyper_synthesis(sd, "h_mean_s", z_hat)
    y_hat, y_lik = slice_loop(sd, y, lm, ls, "forward")
    return {"x_hat": synthesis(sd, y_hat), "likelihoods": {"y": y_lik, "z": z_lik}, "y": y, "z": z, "y_hat": y_hat}


@torch.no_grad()
def compress(sd, x, gc_tab=None, eb_tab=None):
    """:671-732, with per-image y strings (each equals the reference's B=1 string of that image)."""
    gc_tab = gc_tab or entropy.gc_tables()
    ebp = eb_params(sd)
    eb_tab = eb_tab or entropy.eb_tables(ebp)
    y = analysis(sd, x)
    z = conv_stack(sd, "h_a", y, (1, 1, 2, 1, 2))
    z_strings = entropy.eb_compress(ebp, eb_tab, z)
    z_hat = entropy.eb_decompress(ebp, eb_tab, z_strings, z.shape[-2:])
    ls, lm = hyper_synthesis(sd, "h_scale_s", z_hat), hyper_synthesis(sd, "h_mean_s", z_hat)
    _, syms, idxs = slice_loop(sd, y, lm, ls, "compress")
    y_strings = []
    for b in range(x.shape[0]):
        s = np.concatenate([q[b].reshape(-1).numpy() for q in syms])
        i = np.concatenate([q[b].reshape(-1).numpy() for q in idxs])
        y_strings.append(coder.rans_encode(s, i, *gc_tab))
    return {"strings": [y_strings, z_strings], "shape": z.shape[-2:]}


@torch.no_grad()
def decompress(sd, strings, shape, gc_tab=None, eb_tab=None):
    """:734-785 for ONE image (the reference hard-codes B=1, :767); call per image for batches."""
    gc_tab = gc_tab or entropy.gc_tables()
    ebp = eb_params(sd)
    eb_tab = eb_tab or entropy.eb_tables(ebp)
    outs = []
    for b in range(len(strings[1])):
        z_hat = entropy.eb_decompress(ebp, eb_tab, [strings[1][b]], shape)
        ls, lm = hyper_synthesis(sd, "h_scale_s", z_hat), hyper_synthesis(sd, "h_mean_s", z_hat)
        dec = coder.RansDecoder()
        dec.set_stream(strings[0][b])
        y_hat, _, _ = slice_loop(sd, None, lm, ls, "decompress", dec, gc_tab)
        outs.append(synthesis(sd, y_hat).clamp_(0, 1))
    return {"x_hat": torch.cat(outs, 0)}


def template_state_dict():
    """Names, shapes and default values of the reference STF state_dict,
    built without the reference so the seeded weights can be regenerated on any machine."""
    sd = {}
    Z = torch.zeros

    def lin(name, o, i, bias=True):
        sd[name + ".weight"] = Z(o, i)
        if bias:
            sd[name + ".bias"] = Z(o)

    def norm(name, c):
        sd[name + ".weight"] = torch.ones(c)
        sd[name + ".bias"] = Z(c)

    def conv(name, o, i, k):
        sd[name + ".weight"] = Z(o, i, k, k)
        sd[name + ".bias"] = Z(o)

    def layer(prefix, C, depth, heads, down):
        for b in range(depth):
            p = f"{prefix}.blocks.{b}"
            norm(p + ".norm1", C)
            sd[p + ".attn.relative_position_bias_table"] = Z(49, heads)
            sd[p + ".attn.relative_position_index"] = rel_pos_index(WINDOW)
            lin(p + ".attn.qkv", 3 * C, C)
            lin(p + ".attn.proj", C, C)
            norm(p + ".norm2", C)
            lin(p + ".mlp.fc1", 4 * C, C)
            lin(p + ".mlp.fc2", C, 4 * C)
        if down == "merge":
            lin(prefix + ".downsample.reduction", 2 * C, 4 * C, bias=False)
            norm(prefix + ".downsample.norm", 4 * C)
        elif down == "split":
            lin(prefix + ".downsample.reduction", 2 * C, C, bias=False)
            norm(prefix + ".downsample.norm", C)

    conv("patch_embed.proj", 48, 3, 2)
    norm("patch_embed.norm", 48)
    for i in range(4):
        layer(f"layers.{i}", 48 << i, DEPTHS[i], HEADS[i], "merge" if i < 3 else None)
    for i in range(4):
        layer(f"syn_layers.{i}", 384 >> i, DEPTHS[3 - i], HEADS[3 - i], "split" if i < 3 else None)
    conv("end_conv.0", 192, 48, 5)
    conv("end_conv.2", 3, 48, 3)
    for k, (o, i) in zip((0, 2, 4, 6, 8), ((384, 384), (336, 384), (288, 336), (240, 288), (192, 240))):
        conv(f"h_a.{k}", o, i, 3)
    for h in ("h_mean_s", "h_scale_s"):
        conv(f"{h}.0", 240, 192, 3)
        conv(f"{h}.2.0", 1152, 240, 3)
        conv(f"{h}.4", 336, 288, 3)
        conv(f"{h}.6.0", 1536, 336, 3)
        conv(f"{h}.8", 384, 384, 3)
    chans = (224, 176, 128, 64, 32)
    for fam, extra in (("cc_mean_transforms", 0), ("cc_scale_transforms", 0), ("lrp_transforms", 1)):
        for s in range(NUM_SLICES):
            cin = 384 + 32 * min(s + extra, MAX_SUPPORT + extra)
            for k, o in enumerate(chans):
                conv(f"{fam}.{s}.{2 * k}", o, cin, 3)
                cin = o
    f = (1, 3, 3, 3, 3, 1)
    scale = 10 ** (1 / 5)
    for i in range(5):
        sd[f"entropy_bottleneck._matrix{i}"] = torch.full((192, f[i + 1], f[i]), float(np.log(np.expm1(1 / scale / f[i + 1]))))
        sd[f"entropy_bottleneck._bias{i}"] = Z(192, f[i + 1], 1)
        if i < 4:
            sd[f"entropy_bottleneck._factor{i}"] = Z(192, f[i + 1], 1)
    sd["entropy_bottleneck.quantiles"] = torch.tensor([-10.0, 0.0, 10.0]).repeat(192, 1, 1)
    return sd
