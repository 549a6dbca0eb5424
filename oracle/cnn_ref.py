"""Functional fp32 CPU restatement of the reference WACNN transforms + slice loop (row T11).

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Works straight from a reference-named state_dict; pinned against the
real reference module by oracle/make_golden.py (tests/golden/cnn_small.npz) and tests/test_oracle_pinned.py.
Line numbers refer to /root/reference/compressai/models/cnn.py unless another file is named.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import coder, entropy
from .stf_ref import _sub, conv_stack, eb_params, hyper_synthesis, rel_pos_index, slice_loop

NUM_SLICES = 10
MAX_SUPPORT = 5
PEDESTAL = 2.0 ** -36  # ops/parametrizers.py:32-39: reparam_offset ** 2


def nonneg(x, minimum):
    """ops/parametrizers.py:46-49: lower-bound at sqrt(minimum + pedestal), square, subtract pedestal."""
    bound = (minimum + PEDESTAL) ** 0.5
    return torch.clamp(x, min=bound) ** 2 - PEDESTAL


def gdn(x, sd, prefix, inverse):
    """layers/gdn.py:62-75."""
    C = x.shape[1]
    beta = nonneg(sd[prefix + ".beta"], 1e-6)
    gamma = nonneg(sd[prefix + ".gamma"], 0.0).reshape(C, C, 1, 1)
    norm = F.conv2d(x * x, gamma, beta)
    return x * (torch.sqrt(norm) if inverse else torch.rsqrt(norm))


def residual_unit(x, p):
    """layers/layers.py:53-71: 1x1, GELU, 3x3, GELU, 1x1, + identity, GELU."""
    t = F.gelu(F.conv2d(x, p["conv.0.weight"], p["conv.0.bias"]))
    t = F.gelu(F.conv2d(t, p["conv.2.weight"], p["conv.2.bias"], padding=1))
    t = F.conv2d(t, p["conv.4.weight"], p["conv.4.bias"])
    return F.gelu(t + x)


def shifted_window_attention(x, p, heads, win, shift):
    """layers/win_attention.py:147-199 (block) and :86-116 (attention); x [B,C,H,W], no padding."""
    B, C, H, W = x.shape
    hd = C // heads
    t = x.permute(0, 2, 3, 1)
    mask = None
    if shift > 0:
        lab = torch.zeros(H, W)
        edges_h = ((0, H - win), (H - win, H - shift), (H - shift, H))
        edges_w = ((0, W - win), (W - win, W - shift), (W - shift, W))
        for r, (h0, h1) in enumerate(edges_h):
            for c, (w0, w1) in enumerate(edges_w):
                lab[h0:h1, w0:w1] = 3 * r + c
        lw = lab.reshape(H // win, win, W // win, win).permute(0, 2, 1, 3).reshape(-1, win * win)
        diff = lw[:, None, :] - lw[:, :, None]
        mask = torch.where(diff != 0, torch.tensor(-100.0), torch.tensor(0.0))
        t = torch.roll(t, (-shift, -shift), (1, 2))
    xw = t.reshape(B, H // win, win, W // win, win, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, win * win, C)
    Bw, N, _ = xw.shape
    qkv = F.linear(xw, p["attn.qkv.weight"], p["attn.qkv.bias"]).reshape(Bw, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    a = (qkv[0] * hd ** -0.5) @ qkv[1].transpose(-2, -1)
    bias = p["attn.relative_position_bias_table"][rel_pos_index(win).reshape(-1)].reshape(N, N, heads).permute(2, 0, 1)
    a = a + bias.unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        a = (a.reshape(Bw // nW, nW, heads, N, N) + mask[None, :, None]).reshape(-1, heads, N, N)
    o = (torch.softmax(a, -1) @ qkv[2]).transpose(1, 2).reshape(Bw, N, C)
    o = F.linear(o, p["attn.proj.weight"], p["attn.proj.bias"])
    t = o.reshape(B, H // win, W // win, win, win, C).permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, C)
    if shift > 0:
        t = torch.roll(t, (shift, shift), (1, 2))
    return x + t.permute(0, 3, 1, 2)


def gated_window_block(x, sd, prefix, heads, win, shift):
    """layers/layers.py:44-89 (Win_noShift_Attention): x + conv_a(x) * sigmoid(conv_b(x))."""
    a = x
    for k in range(3):
        a = residual_unit(a, _sub(sd, f"{prefix}.conv_a.{k}."))
    b = shifted_window_attention(x, _sub(sd, f"{prefix}.conv_b.0."), heads, win, shift)
    for k in (1, 2, 3):
        b = residual_unit(b, _sub(sd, f"{prefix}.conv_b.{k}."))
    b = F.conv2d(b, sd[f"{prefix}.conv_b.4.weight"], sd[f"{prefix}.conv_b.4.bias"])
    return a * torch.sigmoid(b) + x


def analysis(sd, x):
    """g_a (:31-41)."""
    c = lambda t, n: F.conv2d(t, sd[f"g_a.{n}.weight"], sd[f"g_a.{n}.bias"], stride=2, padding=2)
    t = gdn(c(x, 0), sd, "g_a.1", False)
    t = gdn(c(t, 2), sd, "g_a.3", False)
    t = gated_window_block(t, sd, "g_a.4", 8, 8, 4)
    t = gdn(c(t, 5), sd, "g_a.6", False)
    return gated_window_block(c(t, 7), sd, "g_a.8", 8, 4, 2)


def synthesis(sd, y_hat):
    """g_s (:42-52), unclamped."""
    d = lambda t, n: F.conv_transpose2d(t, sd[f"g_s.{n}.weight"], sd[f"g_s.{n}.bias"], stride=2, padding=2, output_padding=1)
    t = gated_window_block(y_hat, sd, "g_s.0", 8, 4, 2)
    t = gdn(d(t, 1), sd, "g_s.2", True)
    t = gdn(d(t, 3), sd, "g_s.4", True)
    t = gated_window_block(t, sd, "g_s.5", 8, 8, 4)
    t = gdn(d(t, 6), sd, "g_s.7", True)
    return d(t, 8)


def _loop(sd, y, lm, ls, mode, dec=None, tab=None):
    return slice_loop(sd, y, lm, ls, mode, dec, tab, num_slices=NUM_SLICES, max_support=MAX_SUPPORT)


@torch.no_grad()
def forward(sd, x):
    """:141-189 in eval mode."""
    y = analysis(sd, x)
    z = conv_stack(sd, "h_a", y, (1, 1, 2, 1, 2))
    z_hat, z_lik = entropy.eb_forward_eval(eb_params(sd), z)
    ls, lm = hyper_synthesis(sd, "h_scale_s", z_hat), hyper_synthesis(sd, "h_mean_s", z_hat)
    y_hat, y_lik = _loop(sd, y, lm, ls, "forward")
    return {"x_hat": synthesis(sd, y_hat), "likelihoods": {"y": y_lik, "z": z_lik}, "y": y, "z": z, "y_hat": y_hat}


@torch.no_grad()
def compress(sd, x, gc_tab=None, eb_tab=None):
    """:210-268, with per-image y strings (each equals the reference's B=1 string of that image)."""
    gc_tab = gc_tab or entropy.gc_tables()
    ebp = eb_params(sd)
    eb_tab = eb_tab or entropy.eb_tables(ebp)
    y = analysis(sd, x)
    z = conv_stack(sd, "h_a", y, (1, 1, 2, 1, 2))
    z_strings = entropy.eb_compress(ebp, eb_tab, z)
    z_hat = entropy.eb_decompress(ebp, eb_tab, z_strings, z.shape[-2:])
    ls, lm = hyper_synthesis(sd, "h_scale_s", z_hat), hyper_synthesis(sd, "h_mean_s", z_hat)
    _, syms, idxs = _loop(sd, y, lm, ls, "compress")
    y_strings = []
    for b in range(x.shape[0]):
        s = np.concatenate([q[b].reshape(-1).numpy() for q in syms])
        i = np.concatenate([q[b].reshape(-1).numpy() for q in idxs])
        y_strings.append(coder.rans_encode(s, i, *gc_tab))
    return {"strings": [y_strings, z_strings], "shape": z.shape[-2:]}


@torch.no_grad()
def decompress(sd, strings, shape, gc_tab=None, eb_tab=None):
    """:291-332 for ONE image at a time (the reference hard-codes B=1, :316)."""
    gc_tab = gc_tab or entropy.gc_tables()
    ebp = eb_params(sd)
    eb_tab = eb_tab or entropy.eb_tables(ebp)
    outs = []
    for b in range(len(strings[1])):
        z_hat = entropy.eb_decompress(ebp, eb_tab, [strings[1][b]], shape)
        ls, lm = hyper_synthesis(sd, "h_scale_s", z_hat), hyper_synthesis(sd, "h_mean_s", z_hat)
        dec = coder.RansDecoder()
        dec.set_stream(strings[0][b])
        y_hat, _, _ = _loop(sd, None, lm, ls, "decompress", dec, gc_tab)
        outs.append(synthesis(sd, y_hat).clamp_(0, 1))
    return {"x_hat": torch.cat(outs, 0)}


def template_state_dict():
    """Names, shapes and default values of the reference WACNN state_dict, built without the reference."""
    sd = {}
    Z = torch.zeros

    def conv(name, o, i, k):
        sd[name + ".weight"] = Z(o, i, k, k)
        sd[name + ".bias"] = Z(o)

    def deconv(name, i, o, k):
        sd[name + ".weight"] = Z(i, o, k, k)
        sd[name + ".bias"] = Z(o)

    def gdn_(name, c):
        sd[name + ".beta"] = torch.sqrt(torch.ones(c) + PEDESTAL)
        sd[name + ".gamma"] = torch.sqrt(0.1 * torch.eye(c) + PEDESTAL)
        sd[name + ".beta_reparam.pedestal"] = torch.tensor([PEDESTAL])
        sd[name + ".beta_reparam.lower_bound.bound"] = torch.tensor([(1e-6 + PEDESTAL) ** 0.5])
        sd[name + ".gamma_reparam.pedestal"] = torch.tensor([PEDESTAL])
        sd[name + ".gamma_reparam.lower_bound.bound"] = torch.tensor([PEDESTAL ** 0.5])

    def ru(name, c):
        conv(name + ".conv.0", c // 2, c, 1)
        conv(name + ".conv.2", c // 2, c // 2, 3)
        conv(name + ".conv.4", c, c // 2, 1)

    def block(name, c, heads, win):
        for k in range(3):
            ru(f"{name}.conv_a.{k}", c)
        a = f"{name}.conv_b.0.attn"
        sd[a + ".relative_position_bias_table"] = Z((2 * win - 1) ** 2, heads)
        sd[a + ".relative_position_index"] = rel_pos_index(win)
        sd[a + ".qkv.weight"], sd[a + ".qkv.bias"] = Z(3 * c, c), Z(3 * c)
        sd[a + ".proj.weight"], sd[a + ".proj.bias"] = Z(c, c), Z(c)
        for k in (1, 2, 3):
            ru(f"{name}.conv_b.{k}", c)
        conv(f"{name}.conv_b.4", c, c, 1)

    N, M = 192, 320
    conv("g_a.0", N, 3, 5); gdn_("g_a.1", N); conv("g_a.2", N, N, 5); gdn_("g_a.3", N); block("g_a.4", N, 8, 8)
    conv("g_a.5", N, N, 5); gdn_("g_a.6", N); conv("g_a.7", M, N, 5); block("g_a.8", M, 8, 4)
    block("g_s.0", M, 8, 4); deconv("g_s.1", M, N, 5); gdn_("g_s.2", N); deconv("g_s.3", N, N, 5); gdn_("g_s.4", N)
    block("g_s.5", N, 8, 8); deconv("g_s.6", N, N, 5); gdn_("g_s.7", N); deconv("g_s.8", N, 3, 5)
    for k, (o, i) in enumerate(((320, 320), (288, 320), (256, 288), (224, 256), (192, 224))):
        conv(f"h_a.{2 * k}", o, i, 3)
    for pre in ("h_mean_s", "h_scale_s"):
        conv(pre + ".0", 192, 192, 3); conv(pre + ".2.0", 224 * 4, 192, 3); conv(pre + ".4", 256, 224, 3)
        conv(pre + ".6.0", 288 * 4, 256, 3); conv(pre + ".8", 320, 288, 3)
    chans = (224, 176, 128, 64, 32)
    for i in range(NUM_SLICES):
        for pre, cin in (("cc_mean_transforms", 320 + 32 * min(i, 5)), ("cc_scale_transforms", 320 + 32 * min(i, 5)),
                         ("lrp_transforms", 320 + 32 * min(i + 1, 6))):
            for k, o in enumerate(chans):
                conv(f"{pre}.{i}.{2 * k}", o, cin, 3)
                cin = o
    f = (1, 3, 3, 3, 3, 1)
    scale = 10 ** (1 / 5)
    for i in range(5):
        sd[f"entropy_bottleneck._matrix{i}"] = torch.full((N, f[i + 1], f[i]), float(np.log(np.expm1(1 / scale / f[i + 1]))))
        sd[f"entropy_bottleneck._bias{i}"] = Z(N, f[i + 1], 1)
        if i < 4:
            sd[f"entropy_bottleneck._factor{i}"] = Z(N, f[i + 1], 1)
    sd["entropy_bottleneck.quantiles"] = torch.tensor([-10.0, 0.0, 10.0]).repeat(N, 1, 1)
    return sd
