"""CPU restatement of the reference's entropy-model step (rows E1-E9 of SURVEY.md §8a).

TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/rans_oracle.c header for the import rule).
Floating-point parts use torch fp32 on the CPU, because that is what the reference computes with;
integer parts are numpy.  Every function cites the reference lines it follows (paths relative to
/root/reference/compressai/entropy_models/entropy_models.py unless noted).  Pinned against the real
reference by oracle/make_golden.py -> tests/golden/entropy_kat.npz, stf_small.npz.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import coder

SCALE_BOUND = 0.11
LIKELIHOOD_BOUND = 1e-9
TAIL_MASS = 1e-9
# -scipy.stats.norm.ppf(1e-9 / 2) (entropy_models.py:586,600); constant so scipy is not needed on the box
GAUSS_MULTIPLIER = 6.109410204869

def scale_table(lo=0.11, hi=256.0, levels=64):
    """models/stf.py:16-22."""
    return torch.exp(torch.linspace(math.log(lo), math.log(hi), levels))


# ---- E1/E2 ------------------------------------------------------------------------------------
def quantize_symbols(x, means=None):
    """:126-150, mode "symbols": round-half-even of (x - means) as int32."""
    v = x if means is None else x - means
    return torch.round(v).int()


def quantize_dequantize(x, means=None):
    """:126-146, mode "dequantize"."""
    if means is None:
        return torch.round(x)
    return torch.round(x - means) + means


def dequantize(symbols, means=None):
    """:159-165."""
    return symbols.float() if means is None else symbols.type_as(means) + means


# ---- E3 -----------------------------------------------------------------------------------------
def build_indexes(scales, table):
    """:661-666.  idx = (len-1) - #{j < len-1 : max(scale, bound) <= table[j]}."""
    s = torch.clamp_min(scales, SCALE_BOUND)
    idx = torch.full(s.shape, len(table) - 1, dtype=torch.int32)
    for t in table[:-1]:
        idx -= (s <= t).int()
    return idx


# ---- E4 -----------------------------------------------------------------------------------------
def _std_cumulative(v):
    """:578-582."""
    return 0.5 * torch.erfc(float(-(2 ** -0.5)) * v)


def gc_likelihood(x_tilde, scales, means=None):
    """:626-643 + likelihood lower bound (:656-657)."""
    v = x_tilde if means is None else x_tilde - means
    s = torch.clamp_min(scales, SCALE_BOUND)
    v = torch.abs(v)
    p = _std_cumulative((0.5 - v) / s) - _std_cumulative((-0.5 - v) / s)
    return torch.clamp_min(p, LIKELIHOOD_BOUND)


def gc_forward_eval(y, scales, means):
    """GaussianConditional.forward in eval mode (:645-659) -> (y_hat, likelihood)."""
    y_hat = quantize_dequantize(y, means)
    return y_hat, gc_likelihood(y_hat, scales, means)


# ---- E5 -----------------------------------------------------------------------------------------
def eb_logits_cumulative(p, v):
    """:400-419.  p: dict with _matrix{k} [C,fo,fi], _bias{k} [C,fo,1], _factor{k} [C,fo,1]; v [C,1,N]."""
    h = v
    for k in range(5):
        h = torch.matmul(F.softplus(p[f"_matrix{k}"]), h)
        h = h + p[f"_bias{k}"]
        if k < 4:
            h = h + torch.tanh(p[f"_factor{k}"]) * torch.tanh(h)
    return h


def eb_likelihood_cn(p, v):
    """:422-433 on channel-major values v [C,1,N] (+ bound, :476-477)."""
    lo = eb_logits_cumulative(p, v - 0.5)
    up = eb_logits_cumulative(p, v + 0.5)
    sign = -torch.sign(lo + up)
    lik = torch.abs(torch.sigmoid(sign * up) - torch.sigmoid(sign * lo))
    return torch.clamp_min(lik, LIKELIHOOD_BOUND)


def eb_forward_eval(p, z):
    """EntropyBottleneck.forward in eval mode (:446-489) on z [B,C,H,W] -> (z_hat, likelihood)."""
    B, C = z.shape[:2]
    med = p["quantiles"][:, :, 1:2]  # [C,1,1]
    v = z.transpose(0, 1).reshape(C, 1, -1)
    v_hat = torch.round(v - med) + med
    lik = eb_likelihood_cn(p, v_hat)
    back = lambda t: t.reshape(C, B, *z.shape[2:]).transpose(0, 1).contiguous()
    return back(v_hat), back(lik)


# ---- E8 -----------------------------------------------------------------------------------------
def _pmf_rows_to_cdf(pmf, tail, lengths, max_length):
    """:172-180."""
    out = np.zeros((len(lengths), max_length + 2), np.int32)
    for i in range(len(lengths)):
        row = torch.cat((pmf[i, : int(lengths[i])], tail[i]), 0).numpy()
        c = coder.pmf_to_quantized_cdf(row, 16)
        out[i, : c.size] = c
    return out


def gc_tables(table=None):
    """GaussianConditional.update (:599-624) -> (cdf int32 [T, L+2], cdf_length, offset)."""
    table = scale_table() if table is None else table
    mult = -_std_quantile(TAIL_MASS / 2)
    center = torch.ceil(table * mult).int()
    length = 2 * center + 1
    max_length = int(length.max())
    samples = torch.abs(torch.arange(max_length).int() - center[:, None]).float()
    s = table.unsqueeze(1).float()
    upper = _std_cumulative((0.5 - samples) / s)
    lower = _std_cumulative((-0.5 - samples) / s)
    pmf = upper - lower
    tail = 2 * lower[:, :1]
    cdf = _pmf_rows_to_cdf(pmf, tail, length, max_length)
    return cdf, (length + 2).numpy().astype(np.int32), (-center).numpy().astype(np.int32)


def _std_quantile(q):
    """:584-586 (scipy.stats.norm.ppf); falls back to the recorded constant when scipy is missing."""
    try:
        import scipy.stats

        return float(scipy.stats.norm.ppf(q))
    except Exception:  # pragma: no cover
        assert abs(q - TAIL_MASS / 2) < 1e-24
        return -GAUSS_MULTIPLIER


def eb_tables(p):
    """EntropyBottleneck.update (:354-393)."""
    q = p["quantiles"]
    med = q[:, 0, 1]
    minima = torch.clamp(torch.ceil(med - q[:, 0, 0]).int(), min=0)
    maxima = torch.clamp(torch.ceil(q[:, 0, 2] - med).int(), min=0)
    offset = -minima
    start = med - minima
    length = maxima + minima + 1
    max_length = int(length.max())
    samples = torch.arange(max_length)[None, :] + start[:, None, None]
    lower = eb_logits_cumulative(p, samples - 0.5)
    upper = eb_logits_cumulative(p, samples + 0.5)
    sign = -torch.sign(lower + upper)
    pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))[:, 0, :]
    tail = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
    cdf = _pmf_rows_to_cdf(pmf, tail, length, max_length)
    return cdf, (length + 2).numpy().astype(np.int32), offset.numpy().astype(np.int32)


# ---- E6/E7 --------------------------------------------------------------------------------------
def eb_compress(p, tables, z):
    """:508-515 -> :203-238.  One string per image, symbols in (c,h,w) order, index = channel."""
    cdf, lengths, offsets = tables
    B, C = z.shape[:2]
    med = p["quantiles"][:, 0, 1].reshape(1, C, *([1] * (z.dim() - 2)))
    sym = quantize_symbols(z, med)
    idx = torch.arange(C, dtype=torch.int32).reshape(1, C, *([1] * (z.dim() - 2))).expand_as(sym)
    return [coder.rans_encode(sym[b].reshape(-1).numpy(), idx[b].reshape(-1).numpy(), cdf, lengths, offsets) for b in range(B)]


def eb_decompress(p, tables, strings, size):
    """:517-522 -> :240-290."""
    cdf, lengths, offsets = tables
    C = cdf.shape[0]
    med = p["quantiles"][:, 0, 1].reshape(1, C, *([1] * len(size)))
    idx = torch.arange(C, dtype=torch.int32).reshape(C, *([1] * len(size))).expand(C, *size).reshape(-1).numpy()
    out = torch.empty(len(strings), C, *size, dtype=torch.int32)
    for b, s in enumerate(strings):
        v = coder.RansDecoder().decode_with_indexes(s, idx, cdf, lengths, offsets)
        out[b] = torch.from_numpy(v).reshape(C, *size)
    return dequantize(out, med.expand(len(strings), C, *size))
