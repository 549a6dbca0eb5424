"""Training-mode forward of the STF codec (reference models/stf.py:582-645 in `.train()`).

The inference path is hand-written CUDA end to end; the training step keeps the transforms differentiable by running
them as PyTorch operators (library GEMMs / convolutions under autograd -- the backward of the transforms is the host
framework's, as the round-1 verdict allows for this row) and replaces the memory-bound chains around them with the
fused kernels of csrc/train.cu:

  * GaussianConditional in training mode + the straight-through y_hat: `_GaussianTrain` (icm_gc_train_forward /
    icm_gc_train_backward), one pass each way per slice instead of ~20 autograd nodes;
  * gradient-norm clipping + both Adams over flat buffers, and the bucketed gradient all-reduce: compressai/training.py.

Randomness follows the reference exactly in kind and order (DropPath masks per residual branch, stf.py:194-198 with
drop_path_rate 0.2 and dpr = linspace(0, rate, 12), :441,:458-476; uniform noise on z then on each y slice,
entropy_models.py:131-135) through a `TrainRng`, so that a test can replay the reference's CPU draws.
"""
import torch
import torch.nn.functional as F

from compressai._native import NativeError, Rows, check, lib, stream_ptr
from compressai.ops import ste_round

DROP_PATH_RATE = 0.2  # SymmetricalTransFormer(drop_path_rate=0.2), stf.py:398


class TrainRng:
    """Source of the training-mode randomness; draws on the tensor's own device."""

    def uniform(self, like):
        """U(-1/2, 1/2) noise shaped like `like` (entropy_models.py:131-135)."""
        return torch.empty_like(like).uniform_(-0.5, 0.5)

    def keep_mask(self, like, keep):
        """Bernoulli(keep) per sample, [B, 1, 1] (timm DropPath, used at stf.py:194-198)."""
        return like.new_empty((like.shape[0], 1, 1)).bernoulli_(keep)


class ReplayRng(TrainRng):
    """Draws from the CPU generator in the reference's order and shapes, then moves to the tensor's device: with the
    same torch.manual_seed this reproduces a CPU run of the reference modules draw for draw (tests)."""

    def uniform(self, like):
        return torch.empty(like.shape, dtype=torch.float32).uniform_(-0.5, 0.5).to(like.device)

    def keep_mask(self, like, keep):
        return torch.empty((like.shape[0], 1, 1), dtype=torch.float32).bernoulli_(keep).to(like.device)


class _GaussianTrain(torch.autograd.Function):
    """(y, mu, scale, noise) -> (likelihood, y_hat) for one 32-channel slice, fused (csrc/train.cu)."""

    @staticmethod
    def forward(ctx, y, mu, scale, noise, scale_bound, lik_bound):
        if not y.is_cuda:
            raise NativeError("the fused training kernels run on CUDA only (no CPU fallback)")
        if y.dtype != torch.float32:
            y = y.float()
        B = y.shape[0]
        n = y[0].numel()
        mu, scale, noise = mu.contiguous(), scale.contiguous(), noise.contiguous()
        if y.stride(0) % 4 or y[0].is_contiguous() is False or n % 4:
            y = y.contiguous()
        lik, y_hat = torch.empty_like(mu), torch.empty_like(mu)
        r = lambda t: Rows(t.data_ptr(), t.stride(0))
        check(lib().icm_gc_train_forward(r(y), r(noise), r(mu), r(scale), B, n, scale_bound, lik_bound, r(lik), r(y_hat), stream_ptr()),
              "icm_gc_train_forward")
        ctx.save_for_backward(y, mu, scale, noise)
        ctx.bounds = (scale_bound, lik_bound)
        return lik, y_hat

    @staticmethod
    def backward(ctx, g_lik, g_hat):
        y, mu, scale, noise = ctx.saved_tensors
        B, n = y.shape[0], y[0].numel()
        g_lik = (torch.zeros_like(mu) if g_lik is None else g_lik).contiguous()
        g_hat = (torch.zeros_like(mu) if g_hat is None else g_hat).contiguous()
        g_y, g_mu, g_s = torch.empty_like(mu), torch.empty_like(mu), torch.empty_like(mu)
        r = lambda t: Rows(t.data_ptr(), t.stride(0))
        check(lib().icm_gc_train_backward(r(y), r(noise), r(mu), r(scale), r(g_lik), r(g_hat), B, n, ctx.bounds[0], ctx.bounds[1],
                                          r(g_y), r(g_mu), r(g_s), stream_ptr()), "icm_gc_train_backward")
        return g_y, g_mu, g_s, None, None, None


class _LayerNormTrain(torch.autograd.Function):
    """nn.LayerNorm forward / backward as one kernel each (csrc/train.cu) instead of torch's three; emits bf16 directly under
    autocast (the consumer is always a linear layer)."""

    @staticmethod
    def forward(ctx, x, weight, bias, out_bf16):
        C = x.shape[-1]
        x2 = x.reshape(-1, C)
        if x2.dtype != torch.float32 or not x2.is_contiguous():
            x2 = x2.float().contiguous()
        rows = x2.shape[0]
        w, b = weight.detach().float().contiguous(), bias.detach().float().contiguous()
        y = torch.empty((rows, C), dtype=torch.bfloat16 if out_bf16 else torch.float32, device=x.device)
        mean, rstd = torch.empty(rows, dtype=torch.float32, device=x.device), torch.empty(rows, dtype=torch.float32, device=x.device)
        check(lib().icm_layernorm_train_forward(x2.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), 0 if out_bf16 else 1, mean.data_ptr(),
                                                rstd.data_ptr(), rows, C, stream_ptr()), "icm_layernorm_train_forward")
        ctx.save_for_backward(x2, w, mean, rstd)
        ctx.shape = x.shape
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, g):
        x2, w, mean, rstd = ctx.saved_tensors
        rows, C = x2.shape
        g2 = g.reshape(rows, C)
        if g2.dtype not in (torch.float32, torch.bfloat16):
            g2 = g2.float()
        g2 = g2.contiguous()
        dx = torch.empty_like(x2)
        dg, db = torch.empty(C, dtype=torch.float32, device=x2.device), torch.empty(C, dtype=torch.float32, device=x2.device)
        check(lib().icm_layernorm_train_backward(x2.data_ptr(), g2.data_ptr(), 0 if g2.dtype == torch.bfloat16 else 1, w.data_ptr(), mean.data_ptr(),
                                                 rstd.data_ptr(), dx.data_ptr(), dg.data_ptr(), db.data_ptr(), rows, C, stream_ptr()),
              "icm_layernorm_train_backward")
        return dx.view(ctx.shape), dg, db, None


def layer_norm(x, norm, fused=True, fp32_out=False):
    """LayerNorm over the last dimension with the module's parameters; fused kernels on CUDA, the PyTorch operator otherwise
    (fused=False: the comparison the tests make).  Under autocast the fused form emits bf16 (its consumer is a linear layer)
    unless fp32_out (the patch embedding's LayerNorm starts the fp32 residual stream)."""
    C = x.shape[-1]
    if fused and x.is_cuda and C % 4 == 0 and C <= 768:
        return _LayerNormTrain.apply(x, norm.weight, norm.bias, torch.is_autocast_enabled() and not fp32_out)
    return F.layer_norm(x, (C,), norm.weight, norm.bias)


def gaussian_train(gc, y, mu, scale, noise):
    """GaussianConditional(y, scale, mu) in training mode -> (likelihood, straight-through y_hat)."""
    lb = gc.likelihood_bound if gc.use_likelihood_bound else 0.0
    return _GaussianTrain.apply(y, mu, scale, noise, float(gc._scale_bound_f), float(lb))


# ---------------------------------------------------------------------------------------------------- Swin pieces
def _shift_mask(H, W, ws, shift, device):
    """[nW, ws*ws, ws*ws] additive mask of the shifted windows (stf.py:316-334): tokens of different wrap-around regions
    do not attend to each other (-100)."""
    Hp, Wp = -(-H // ws) * ws, -(-W // ws) * ws
    def labels(n):
        t = torch.zeros(n, dtype=torch.float32, device=device)
        t[n - ws:n - shift] = 1
        t[n - shift:] = 2
        return t
    lab = labels(Hp)[:, None] * 3 + labels(Wp)[None, :]
    win = lab.view(Hp // ws, ws, Wp // ws, ws).permute(0, 2, 1, 3).reshape(-1, ws * ws)
    diff = win[:, None, :] - win[:, :, None]
    return torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))


def _window_attention(attn, x, H, W, ws, shift, mask):
    """x [B, H, W, C] (LayerNorm-ed) -> attention branch output [B, H, W, C] (stf.py:90-121, 157-191)."""
    B, _, _, C = x.shape
    pad_b, pad_r = (-H) % ws, (-W) % ws
    if pad_b or pad_r:
        x = F.pad(x, (0, 0, 0, pad_r, 0, pad_b))
    Hp, Wp = H + pad_b, W + pad_r
    if shift:
        x = torch.roll(x, (-shift, -shift), (1, 2))
    nh = attn.num_heads
    hd = C // nh
    win = x.view(B, Hp // ws, ws, Wp // ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws, C)  # [B*nW, N, C]
    qkv = F.linear(win, attn.qkv.weight, attn.qkv.bias).view(-1, ws * ws, 3, nh, hd).permute(2, 0, 3, 1, 4)
    bias = attn.relative_position_bias_table[attn.relative_position_index.reshape(-1)].view(ws * ws, ws * ws, nh).permute(2, 0, 1)
    add = bias[None]                                                  # [1, nh, N, N]
    if shift:
        nW = mask.shape[0]
        add = (add[:, None] + mask[None, :, None]).expand(B, nW, nh, ws * ws, ws * ws).reshape(-1, nh, ws * ws, ws * ws)
    out = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2], attn_mask=add.to(qkv.dtype), scale=hd ** -0.5)
    out = F.linear(out.transpose(1, 2).reshape(-1, ws * ws, C), attn.proj.weight, attn.proj.bias)
    x = out.view(B, Hp // ws, Wp // ws, ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, C)
    if shift:
        x = torch.roll(x, (shift, shift), (1, 2))
    return x[:, :H, :W, :] if (pad_b or pad_r) else x


def _drop_path(t, rate, rng):
    if rate == 0.0:
        return t
    keep = 1.0 - rate
    return t * rng.keep_mask(t, keep).to(t.dtype) / keep


def _stage(layer, x, H, W, rates, rng, fused=True):
    """BasicLayer (stf.py:308-347): x [B, H*W, C] tokens."""
    B, _, C = x.shape
    ws = layer.blocks[0].window_size
    mask = _shift_mask(H, W, ws, ws // 2, x.device) if len(layer.blocks) > 1 else None
    for blk, rate in zip(layer.blocks, rates):
        a = _window_attention(blk.attn, layer_norm(x, blk.norm1, fused).view(B, H, W, C), H, W, ws, blk.shift_size, mask)
        x = x + _drop_path(a.reshape(B, H * W, C), rate, rng)
        h = layer_norm(x, blk.norm2, fused)
        h = F.linear(F.gelu(F.linear(h, blk.mlp.fc1.weight, blk.mlp.fc1.bias)), blk.mlp.fc2.weight, blk.mlp.fc2.bias)
        x = x + _drop_path(h, rate, rng)
    ds = layer.downsample
    if ds is None:
        return x, H, W
    if ds.kind == "merge":  # stf.py:209-235
        g = x.view(B, H, W, C)
        if H % 2 or W % 2:
            g = F.pad(g, (0, 0, 0, W % 2, 0, H % 2))
        g = torch.cat([g[:, 0::2, 0::2], g[:, 1::2, 0::2], g[:, 0::2, 1::2], g[:, 1::2, 1::2]], -1)
        H, W = (H + 1) // 2, (W + 1) // 2
        g = layer_norm(g.reshape(B, H * W, 4 * C), ds.norm, fused)
        return F.linear(g, ds.reduction.weight).float(), H, W  # the residual stream stays fp32 under autocast
    g = F.linear(layer_norm(x, ds.norm, fused), ds.reduction.weight)  # stf.py:251-260
    g = F.pixel_shuffle(g.transpose(1, 2).reshape(B, 2 * C, H, W), 2)
    return g.permute(0, 2, 3, 1).reshape(B, 4 * H * W, C // 2).float(), 2 * H, 2 * W


def _conv_stack(seq, x):
    for m in seq:
        if isinstance(m, torch.nn.Conv2d):
            x = F.conv2d(x, m.weight, m.bias, m.stride, m.padding)
        elif isinstance(m, torch.nn.GELU):
            x = F.gelu(x)
        elif isinstance(m, torch.nn.PixelShuffle):
            x = F.pixel_shuffle(x, m.upscale_factor)
        elif isinstance(m, torch.nn.Sequential):  # subpel_conv3x3: conv + PixelShuffle
            x = F.pixel_shuffle(F.conv2d(x, m[0].weight, m[0].bias, m[0].stride, m[0].padding), m[1].upscale_factor)
        else:
            raise TypeError(type(m))
    return x


def drop_path_rates(depths, rate=DROP_PATH_RATE):
    """Per-block rates of the analysis and synthesis stages (stf.py:441 and the slices at :458, :476: the synthesis
    stages index the SAME ascending list with the reversed depths)."""
    dpr = [float(v) for v in torch.linspace(0, rate, sum(depths))]
    def cut(ds):
        out, o = [], 0
        for d in ds:
            out.append(dpr[o:o + d])
            o += d
        return out
    return cut(depths), cut(depths[::-1])


def stf_train_forward(m, x, rng=None, fused=True):
    """SymmetricalTransFormer.forward in training mode (stf.py:582-645): {"x_hat", "likelihoods": {"y", "z"}} with autograd.
    fused=False evaluates the Gaussian stage with the PyTorch expressions of compressai.entropy_models instead of the
    fused kernels (the comparison the GPU tests make)."""
    rng = rng or TrainRng()
    depths = [len(l.blocks) for l in m.layers]
    ana_rates, syn_rates = drop_path_rates(depths)
    pe = m.patch_embed
    B, _, H, W = x.shape
    if W % 2:
        x = F.pad(x, (0, 1))
    if H % 2:
        x = F.pad(x, (0, 0, 0, 1))
    t = F.conv2d(x, pe.proj.weight, pe.proj.bias, stride=2)
    h, w = t.shape[2], t.shape[3]
    t = layer_norm(t.flatten(2).transpose(1, 2).float(), pe.norm, fused, fp32_out=True).float()
    for layer, rates in zip(m.layers, ana_rates):
        t, h, w = _stage(layer, t, h, w, rates, rng, fused)
    M, Z = m.latent_channels, 32
    y = t.view(B, h, w, M).permute(0, 3, 1, 2).contiguous()
    z = _conv_stack(m.h_a, y)
    eb = m.entropy_bottleneck
    # EntropyBottleneck.forward(training): noise drawn in the [C, 1, B*H*W] layout (entropy_models.py:446-489)
    zv = z.permute(1, 0, 2, 3).reshape(z.shape[1], 1, -1)
    z_lik = eb._likelihood(zv + rng.uniform(zv))
    if eb.use_likelihood_bound:
        z_lik = eb.likelihood_lower_bound(z_lik)
    z_lik = z_lik.view(z.shape[1], B, z.shape[2], z.shape[3]).permute(1, 0, 2, 3).contiguous()
    med = eb._get_medians().view(1, -1, 1, 1)
    z_hat = ste_round(z - med) + med
    scales = _conv_stack(m.h_scale_s, z_hat)
    means = _conv_stack(m.h_mean_s, z_hat)
    gc = m.gaussian_conditional
    hats, liks = [], []
    for i in range(m.num_slices):
        support = hats[:m.max_support_slices]
        mean_sup = torch.cat([means] + support, 1)
        mu = _conv_stack(m.cc_mean_transforms[i], mean_sup)[:, :, :h, :w]
        sc = _conv_stack(m.cc_scale_transforms[i], torch.cat([scales] + support, 1))[:, :, :h, :w]
        ys = y[:, Z * i:Z * (i + 1)]
        noise = rng.uniform(ys)
        if fused:
            lik, y_hat = gaussian_train(gc, ys, mu.float(), sc.float(), noise)
        else:
            lik = _gaussian_torch(gc, ys, mu, sc, noise)
            y_hat = ste_round(ys - mu) + mu
        liks.append(lik)
        lrp = _conv_stack(m.lrp_transforms[i], torch.cat([mean_sup, y_hat], 1))
        hats.append(y_hat + 0.5 * torch.tanh(lrp))
    t = torch.cat(hats, 1).permute(0, 2, 3, 1).reshape(B, h * w, M).float()
    for layer, rates in zip(m.syn_layers, syn_rates):
        t, h, w = _stage(layer, t, h, w, rates, rng, fused)
    u = t.view(B, h, w, m.embed_dim).permute(0, 3, 1, 2).contiguous()
    x_hat = _conv_stack(m.end_conv, u)
    return {"x_hat": x_hat, "likelihoods": {"y": torch.cat(liks, 1), "z": z_lik}}


def _gaussian_torch(gc, ys, mu, sc, noise):
    lik = gc._likelihood(ys + noise, sc, mu)
    return gc.likelihood_lower_bound(lik) if gc.use_likelihood_bound else lik
