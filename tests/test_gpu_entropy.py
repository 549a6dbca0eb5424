"""Entropy-stage kernels (csrc/entropy.cu) through the reference-shaped nn.Module API vs fixtures recorded
from the reference's Python modules and vs the pinned oracle.  Integer outputs: bit-exact.  Likelihoods:
relative 1e-3 (north_star tolerance)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
LIK_RTOL = 1e-3


@pytest.fixture(scope="module")
def gc():
    from compressai.entropy_models import GaussianConditional
    from compressai.models.stf import get_scale_table

    m = GaussianConditional(None).cuda().eval()
    m.update_scale_table(get_scale_table())
    return m


def test_gaussian_conditional_kats(gc, golden_dir):
    g = np.load(os.path.join(golden_dir, "entropy_kat.npz"))
    t = lambda k: torch.from_numpy(g[k]).cuda().reshape(1, 1, -1)
    idx = gc.build_indexes(t("scales"))
    assert idx.dtype == torch.int32 and np.array_equal(idx.cpu().numpy().ravel(), g["indexes"])
    sym = gc.quantize(t("y"), "symbols", t("means"))
    assert sym.dtype == torch.int32 and np.array_equal(sym.cpu().numpy().ravel(), g["symbols"])
    deq = gc.quantize(t("y"), "dequantize", t("means"))
    assert np.array_equal(deq.cpu().numpy().ravel(), g["y_hat"])
    assert np.array_equal(gc.dequantize(sym, t("means")).cpu().numpy().ravel(), g["y_hat"])
    y_hat, lik = gc(t("y"), t("scales"), t("means"))
    assert np.array_equal(y_hat.cpu().numpy().ravel(), g["y_hat"])
    ref = g["y_lik"]
    got = lik.cpu().numpy().ravel()
    assert np.all(np.abs(got - ref) <= LIK_RTOL * ref), float(np.max(np.abs(got - ref) / ref))
    # half-to-even rounding KAT
    q = gc.quantize(torch.tensor([[.5, 1.5, 2.5, -.5, -1.5, -2.5]]).cuda(), "symbols")
    assert q.cpu().tolist() == [[0, 2, 2, 0, -2, -2]]


def test_layouts_agree(gc):
    """The same kernel on NCHW tensors (module API) and on channels-last views (model path)."""
    from compressai._native import NULL_VIEW, check, lib, stream_ptr, view_bcp

    torch.manual_seed(3)
    B, C, H, W = 3, 32, 5, 7
    P = H * W
    y = torch.randn(B, C, H, W, device="cuda") * 6
    mu = torch.randn(B, C, H, W, device="cuda")
    sc = torch.exp(torch.empty(B, C, H, W, device="cuda").uniform_(-3, 5))
    sym = gc.quantize(y, "symbols", mu)
    idx = gc.build_indexes(sc)
    cl = lambda t: t.permute(0, 2, 3, 1).contiguous().reshape(B * P, C)
    wide = torch.zeros(B * P, 96, device="cuda")
    wide[:, 40:72] = cl(y)
    s2 = torch.empty(B, 2 * C * P, dtype=torch.int32, device="cuda")
    i2 = torch.empty_like(s2)
    yh = torch.zeros(B * P, 64, device="cuda")
    bf = torch.zeros(B * P, 48, dtype=torch.bfloat16, device="cuda")
    tab = gc.scale_table_device(y.device)
    mu_cl, sc_cl = cl(mu), cl(sc)
    check(lib().icm_gc_quantize_index(view_bcp(wide, B, C, P, 40), view_bcp(mu_cl, B, C, P), view_bcp(sc_cl, B, C, P), B, C, P,
                                      tab.data_ptr(), tab.numel(), 0.11, s2.data_ptr(), i2.data_ptr(), 2 * C * P, C * P,
                                      view_bcp(yh, B, C, P, 32), view_bcp(bf, B, C, P, 16), NULL_VIEW, stream_ptr()))
    assert torch.equal(s2[:, C * P:].reshape(B, C, H, W), sym)
    assert torch.equal(i2[:, C * P:].reshape(B, C, H, W), idx)
    y_hat = sym.float() + mu
    assert torch.equal(yh[:, 32:], cl(y_hat)) and torch.all(yh[:, :32] == 0)
    assert torch.equal(bf[:, 16:].float(), cl(y_hat).bfloat16().float()) and torch.all(bf[:, :16] == 0)


def _indexes_channels_last(scales_flat, table, bound=0.11):
    """icm_gc_build_indexes on a channels-last [1, P, 32] view (the vectorised kernel's layout) -> flat indexes."""
    from compressai._native import check, lib, stream_ptr, view_bcp

    n = scales_flat.numel()
    P = -(-n // 32)
    sc = torch.full((P * 32,), 1.0, device="cuda")
    sc[:n] = scales_flat
    sc = sc.reshape(P, 32).contiguous()
    idx = torch.empty(1, 32 * P, dtype=torch.int32, device="cuda")
    check(lib().icm_gc_build_indexes(view_bcp(sc, 1, 32, P), 1, 32, P, table.data_ptr(), table.numel(), bound, idx.data_ptr(), 32 * P, 0, stream_ptr()))
    return idx.reshape(32, P).t().reshape(-1)[:n]


def test_vectorised_index_lookup_is_exact_on_every_table_edge(gc, golden_dir):
    """The channels-last kernels resolve build_indexes with a two-probe lookup (bit-pattern bins + one compare) instead of
    the binary search.  It must give the reference's index (entropy_models.py:661-666) on the KAT scales -- every table
    value and its two fp32 neighbours, 0, -1, 0.11, 256, 1e9 -- on inf / NaN, and fall back to the search for tables the
    lookup cannot serve (two levels in one bin, more than 256 bins)."""
    g = np.load(os.path.join(golden_dir, "entropy_kat.npz"))
    tab = gc.scale_table_device(torch.device("cuda"))
    scales = torch.from_numpy(g["scales"]).cuda()
    got = _indexes_channels_last(scales, tab)
    assert np.array_equal(got.cpu().numpy(), g["indexes"])
    special = torch.tensor([float("inf"), float("nan"), -float("inf"), -0.0, 1e-30, 3e38], device="cuda")
    ref = gc.build_indexes(special.reshape(1, 1, -1)).reshape(-1)  # generic kernel (binary search)
    assert torch.equal(_indexes_channels_last(special, tab), ref)
    # without the lower bound negative scales reach the lookup itself
    neg = torch.tensor([-5.0, -0.11, 0.0, 0.05, 0.2, 300.0, float("nan")], device="cuda")
    t = tab.cpu()
    want = torch.tensor([int((t[:-1] < v).sum()) if v == v else len(t) - 1 for v in neg.cpu()], dtype=torch.int32)
    assert torch.equal(_indexes_channels_last(neg, tab, bound=-1e30).cpu(), want)
    # tables the lookup must refuse: levels 1 % apart (several per bin) and a range of > 256 bins
    torch.manual_seed(1)
    s = torch.exp(torch.empty(4000, device="cuda").uniform_(-12, 12))
    for t2 in (torch.exp(torch.linspace(-1.0, 1.0, 200)), torch.exp(torch.linspace(-11.0, 11.0, 40))):
        t2 = t2.float().cuda().contiguous()
        want = torch.tensor([int((t2[:-1] < v).sum()) for v in torch.maximum(s, torch.tensor(0.11, device="cuda"))], dtype=torch.int32)
        assert torch.equal(_indexes_channels_last(s, t2).cpu(), want)


def test_entropy_bottleneck_kats(golden_dir):
    from compressai.entropy_models import EntropyBottleneck
    from oracle import weights

    g = np.load(os.path.join(golden_dir, "entropy_kat.npz"))
    eb = EntropyBottleneck(192)
    sd = weights.seeded_state_dict(dict(eb.state_dict()), seed=3, stress=False)
    eb.load_state_dict({k: v for k, v in sd.items() if k in dict(eb.named_parameters())}, strict=False)
    eb.update(force=True)
    eb = eb.cuda().eval()
    z = torch.from_numpy(g["z"]).cuda()
    z_hat, lik = eb(z)
    assert np.array_equal(z_hat.cpu().numpy(), g["z_hat"])
    ref = g["z_lik"]
    got = lik.cpu().numpy()
    assert np.all(np.abs(got - ref) <= LIK_RTOL * ref), float(np.max(np.abs(got - ref) / ref))
    strings = eb.compress(z)
    assert strings[0] == g["z_string0"].tobytes() and strings[1] == g["z_string1"].tobytes()
    assert np.array_equal(eb.decompress(strings, z.shape[-2:]).cpu().numpy(), g["z_hat"])
    with pytest.raises(ValueError):
        eb.decompress("notalist", (3, 5))
    from compressai.entropy_models import EntropyModel

    with pytest.raises(ValueError):  # strings / indexes batch mismatch
        EntropyModel.decompress(eb, strings[:1], eb._build_indexes((2, 192, 3, 5)).cuda())


def test_reference_stage_inputs_give_reference_bitstring(gc, golden_dir):
    """Identical (y, mu, scale) from the reference run -> identical symbols, indexes and y bit-string
    (north_star: 'bit-exact given identical y_hat/scales/means')."""
    from compressai import ans

    g = np.load(os.path.join(golden_dir, "stf_small.npz"))
    y, mu, sc = (torch.from_numpy(g[k]).cuda() for k in ("y", "mu", "scale"))
    sym = gc.quantize(y, "symbols", mu)
    idx = gc.build_indexes(sc)
    assert np.array_equal(sym.cpu().numpy(), g["symbols"])
    assert np.array_equal(idx.cpu().numpy(), g["indexes"].astype(np.int32))
    order = lambda t: torch.cat([c.reshape(1, -1) for c in t.chunk(12, 1)], 1)  # slice-major, then (c,h,w)
    b = ans.encode_streams(gc.device_tables(), order(sym), order(idx))[0]
    assert b == g["y_string"].tobytes()
    e = ans.BufferedRansEncoder()
    e.encode_with_indexes(order(sym).reshape(-1).tolist(), order(idx).reshape(-1).tolist(), gc.quantized_cdf.tolist(),
                          gc.cdf_length.reshape(-1).int().tolist(), gc.offset.reshape(-1).int().tolist())
    assert e.flush() == b
    y_hat, lik = gc(y, sc, mu)
    ref = g["y_lik"]
    got = lik.cpu().numpy()
    assert np.all(np.abs(got - ref) <= LIK_RTOL * ref)


def test_full_size_properties(gc):
    """BASELINE shape (B=8 x 384 x 48 x 32): idempotence and consistency properties + a sample vs the oracle."""
    from oracle import entropy

    torch.manual_seed(0)
    y = torch.randn(8, 384, 48, 32, device="cuda") * 5
    mu = torch.randn_like(y)
    sc = torch.exp(torch.empty_like(y).uniform_(-3, 5.5))
    sym = gc.quantize(y, "symbols", mu)
    y_hat = gc.dequantize(sym, mu)
    assert torch.equal(gc.quantize(y_hat, "symbols", mu), sym)  # quantise(dequantise(q)) == q
    assert torch.equal(gc.quantize(y, "dequantize", mu), y_hat)
    idx = gc.build_indexes(sc)
    assert int(idx.min()) >= 0 and int(idx.max()) <= 63
    tab = gc.scale_table.cuda()
    s = torch.clamp_min(sc, 0.11)
    lo = torch.where(idx > 0, tab[(idx - 1).clamp_min(0).long()], torch.zeros_like(s))
    assert torch.all((idx == 63) | (tab[idx.long()] >= s)) and torch.all(lo < s)  # bucket edges
    sl = (slice(0, 1), slice(0, 64))
    assert torch.equal(idx[sl].cpu(), entropy.build_indexes(sc[sl].cpu(), entropy.scale_table()))
    assert torch.equal(sym[sl].cpu(), entropy.quantize_symbols(y[sl].cpu(), mu[sl].cpu()))
