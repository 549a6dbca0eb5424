"""The CPU oracle (oracle/) replayed against fixtures recorded from the reference itself.

Fixtures come from oracle/make_golden.py: the reference's shipped rANS/CDF binaries and its unmodified
Python entropy models / STF model.  If these pass, the oracle is pinned and the GPU parity tests may
trust it at sizes the fixtures do not cover.
"""
import hashlib
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import cnn_ref, coder, entropy, stf_ref, weights
from oracle.make_golden import seeded_stream


@pytest.fixture(scope="module")
def kat(golden_dir):
    with open(os.path.join(golden_dir, "rans_kat.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def gc_tab():
    return entropy.gc_tables()


def test_pmf_to_quantized_cdf_kats(kat):
    assert len(kat["pmf"]) >= 10
    for k in kat["pmf"]:
        got = coder.pmf_to_quantized_cdf(np.array(k["pmf"], np.float32), k["precision"])
        assert got.tolist() == k["cdf"]


def test_gaussian_tables_match_reference(kat, gc_tab, golden_dir):
    g = np.load(os.path.join(golden_dir, "gc_tables.npz"))
    cdf, lengths, offsets = gc_tab
    assert np.array_equal(lengths, g["cdf_length"])
    assert np.array_equal(offsets, g["offset"])
    for r in (0, 10, 40, 63):
        assert np.array_equal(cdf[r, : lengths[r]], g[f"row{r}"])
    assert hashlib.sha1(cdf.astype("<i4").tobytes()).hexdigest() == kat["gc_table_sha1"]
    assert abs(float(g["multiplier"]) - entropy.GAUSS_MULTIPLIER) < 1e-9


def test_rans_small_kats(kat):
    t = kat["small_tables"]
    for k in kat["rans_small"]:
        b = coder.rans_encode(k["symbols"], k["indexes"], t["cdfs"], t["sizes"], t["offsets"])
        assert b.hex() == k["hex"]
        d = coder.RansDecoder().decode_with_indexes(b, k["indexes"], t["cdfs"], t["sizes"], t["offsets"])
        assert d.tolist() == k["symbols"]


def test_rans_seeded_streams(kat, gc_tab):
    cdf, lengths, offsets = gc_tab
    table = entropy.scale_table().numpy()
    for k in kat["streams"]:
        kind = "uniform" if k["kind"] == "adversarial" else k["kind"]
        sym, idx = seeded_stream(k["n"], k["seed"], table, kind)
        if k["kind"] == "adversarial":
            n = k["n"]
            c = -offsets[idx]
            a = np.arange(n) % 4
            sym = np.where(a == 0, c, np.where(a == 1, -c, np.where(a == 2, c + 1, -c - 1))).astype(np.int32)
            sym[::97] = 100000
            sym[1::97] = -100000
        assert hashlib.sha1(sym.astype("<i4").tobytes()).hexdigest() == k["sym_sha1"]
        b = coder.rans_encode(sym, idx, cdf, lengths, offsets)
        assert len(b) == k["nbytes"] and hashlib.sha1(b).hexdigest() == k["sha1"]
        # decode in three decode_stream calls on one set_stream (stf.py:751-766 usage)
        d = coder.RansDecoder()
        d.set_stream(b)
        n = k["n"]
        parts = [d.decode_stream(idx[s:e], cdf, lengths, offsets) for s, e in ((0, n // 5), (n // 5, n // 2), (n // 2, n))]
        assert np.array_equal(np.concatenate(parts), sym)
        # buffered multi-call encode == one-shot
        e = coder.BufferedRansEncoder()
        e.encode_with_indexes(sym[: n // 3], idx[: n // 3], cdf, lengths, offsets)
        e.encode_with_indexes(sym[n // 3:], idx[n // 3:], cdf, lengths, offsets)
        assert e.flush() == b


def test_entropy_model_kats(golden_dir):
    g = np.load(os.path.join(golden_dir, "entropy_kat.npz"))
    t = lambda k: torch.from_numpy(g[k])
    table = entropy.scale_table()
    assert np.array_equal(entropy.build_indexes(t("scales"), table).numpy(), g["indexes"])
    # the bucketize form used by the CUDA kernel is the same function
    bz = torch.bucketize(torch.clamp_min(t("scales"), 0.11), table[:-1], right=False).int()
    assert np.array_equal(bz.numpy(), g["indexes"])
    assert np.array_equal(entropy.quantize_symbols(t("y"), t("means")).numpy(), g["symbols"])
    y_hat, lik = entropy.gc_forward_eval(t("y"), t("scales"), t("means"))
    assert torch.equal(y_hat, t("y_hat"))
    assert torch.equal(lik, t("y_lik"))


def test_entropy_bottleneck_kats(golden_dir):
    g = np.load(os.path.join(golden_dir, "entropy_kat.npz"))
    tmpl = {}
    f = (1, 3, 3, 3, 3, 1)
    for i in range(5):
        tmpl[f"_matrix{i}"] = torch.full((192, f[i + 1], f[i]), float(np.log(np.expm1(1 / (10 ** (1 / 5)) / f[i + 1]))))
        tmpl[f"_bias{i}"] = torch.zeros(192, f[i + 1], 1)
        if i < 4:
            tmpl[f"_factor{i}"] = torch.zeros(192, f[i + 1], 1)
    tmpl["quantiles"] = torch.tensor([-10.0, 0, 10]).repeat(192, 1, 1)
    # names inside make_golden were those of a bare EntropyBottleneck module
    p = weights.seeded_state_dict({f"{k}": v for k, v in tmpl.items()}, seed=3, stress=False)
    assert not torch.equal(p["_bias0"], tmpl["_bias0"])
    cdf, lengths, offsets = entropy.eb_tables(p)
    assert np.array_equal(lengths, g["eb_len"]) and np.array_equal(offsets, g["eb_off"])
    assert np.array_equal(cdf, g["eb_cdf"])
    z = torch.from_numpy(g["z"])
    z_hat, z_lik = entropy.eb_forward_eval(p, z)
    assert torch.equal(z_hat, torch.from_numpy(g["z_hat"]))
    assert torch.allclose(z_lik, torch.from_numpy(g["z_lik"]), rtol=1e-6, atol=0)
    strings = entropy.eb_compress(p, (cdf, lengths, offsets), z)
    assert strings[0] == g["z_string0"].tobytes() and strings[1] == g["z_string1"].tobytes()
    assert torch.equal(entropy.eb_decompress(p, (cdf, lengths, offsets), strings, z.shape[-2:]), z_hat)


def test_stf_restatement_matches_reference_model(golden_dir):
    """oracle/stf_ref.py (functional fp32) vs the reference nn.Modules run in make_golden.py."""
    g = np.load(os.path.join(golden_dir, "stf_small.npz"))
    sd = weights.seeded_state_dict(stf_ref.template_state_dict(), seed=0, stress=True)
    x = weights.seeded_image((1, 3, 128, 192), seed=0)
    out = stf_ref.forward(sd, x)
    tol = dict(rtol=1e-4, atol=1e-5)
    assert torch.allclose(out["y"], torch.from_numpy(g["y"]), **tol)
    assert torch.allclose(out["z"], torch.from_numpy(g["z"]), **tol)
    assert torch.allclose(out["x_hat"], torch.from_numpy(g["x_hat"]), rtol=1e-3, atol=1e-4)
    lik = out["likelihoods"]["y"]
    assert torch.allclose(lik, torch.from_numpy(g["y_lik"]), rtol=1e-3, atol=1e-9)
    assert torch.allclose(out["likelihoods"]["z"], torch.from_numpy(g["z_lik"]), rtol=1e-3, atol=1e-9)
    # tables from the seeded EB parameters equal the reference's update()
    eb_tab = entropy.eb_tables(stf_ref.eb_params(sd))
    assert np.array_equal(eb_tab[0], g["eb_cdf"]) and np.array_equal(eb_tab[1], g["eb_len"]) and np.array_equal(eb_tab[2], g["eb_off"])
    # entropy stage on the reference's own (y, mu, scale): symbols / indexes / bitstring bit-exact
    y, mu, sc = (torch.from_numpy(g[k]) for k in ("y", "mu", "scale"))
    sym = entropy.quantize_symbols(y, mu)
    idx = entropy.build_indexes(sc, entropy.scale_table())
    assert np.array_equal(sym.numpy(), g["symbols"]) and np.array_equal(idx.numpy(), g["indexes"].astype(np.int32))
    order = lambda t: np.concatenate([c.reshape(-1).numpy() for c in t[0:1].chunk(12, 1)])
    b = coder.rans_encode(order(sym), order(idx), *entropy.gc_tables())
    assert b == g["y_string"].tobytes()
    # end-to-end through the restatement: same strings, and decompress == clamp(forward)
    c = stf_ref.compress(sd, x, eb_tab=eb_tab)
    assert c["strings"][1][0] == g["z_string"].tobytes()
    assert c["strings"][0][0] == g["y_string"].tobytes()
    d = stf_ref.decompress(sd, c["strings"], c["shape"], eb_tab=eb_tab)
    assert torch.allclose(d["x_hat"], torch.from_numpy(g["x_hat"]).clamp(0, 1), rtol=1e-3, atol=1e-4)


def test_wacnn_restatement_matches_reference_model(golden_dir):
    """oracle/cnn_ref.py (functional fp32) vs the reference WACNN nn.Module run in make_golden.py (256x256, configs[0])."""
    g = np.load(os.path.join(golden_dir, "cnn_small.npz"))
    sd = weights.seeded_state_dict(cnn_ref.template_state_dict(), seed=0, stress=True)
    x = weights.seeded_image((1, 3, 256, 256), seed=0)
    # intermediate activations: first GDN and first gated window block of g_a (stored subsampled)
    t = F.conv2d(x, sd["g_a.0.weight"], sd["g_a.0.bias"], stride=2, padding=2)
    t = cnn_ref.gdn(t, sd, "g_a.1", False)
    assert torch.allclose(t[:, :, ::4, ::4], torch.from_numpy(g["gdn0"]), rtol=1e-4, atol=1e-5)
    out = cnn_ref.forward(sd, x)
    tol = dict(rtol=1e-4, atol=2e-5)
    assert torch.allclose(out["y"], torch.from_numpy(g["y"]), rtol=1e-3, atol=1e-4)
    assert torch.allclose(out["z"], torch.from_numpy(g["z"]), rtol=1e-3, atol=1e-4)
    assert torch.allclose(out["x_hat"], torch.from_numpy(g["x_hat"]), rtol=1e-3, atol=1e-4)
    assert torch.allclose(out["likelihoods"]["y"], torch.from_numpy(g["y_lik"]), rtol=1e-3, atol=1e-9)
    assert torch.allclose(out["likelihoods"]["z"], torch.from_numpy(g["z_lik"]), rtol=1e-3, atol=1e-9)
    eb_tab = entropy.eb_tables(stf_ref.eb_params(sd))
    assert np.array_equal(eb_tab[0], g["eb_cdf"]) and np.array_equal(eb_tab[1], g["eb_len"]) and np.array_equal(eb_tab[2], g["eb_off"])
    y, mu, sc = (torch.from_numpy(g[k]) for k in ("y", "mu", "scale"))
    sym = entropy.quantize_symbols(y, mu)
    idx = entropy.build_indexes(sc, entropy.scale_table())
    assert np.array_equal(sym.numpy(), g["symbols"]) and np.array_equal(idx.numpy(), g["indexes"].astype(np.int32))
    order = lambda t: np.concatenate([c.reshape(-1).numpy() for c in t[0:1].chunk(10, 1)])
    assert coder.rans_encode(order(sym), order(idx), *entropy.gc_tables()) == g["y_string"].tobytes()
    c = cnn_ref.compress(sd, x, eb_tab=eb_tab)
    assert c["strings"][1][0] == g["z_string"].tobytes()
    assert c["strings"][0][0] == g["y_string"].tobytes()
    d = cnn_ref.decompress(sd, c["strings"], c["shape"], eb_tab=eb_tab)
    assert torch.allclose(d["x_hat"], torch.from_numpy(g["x_hat"]).clamp(0, 1), rtol=1e-3, atol=1e-4)


def test_wacnn_restatement_at_config4_size(golden_dir):
    """The 832x1216 WACNN2-codec case (BASELINE.json configs[3]): oracle/cnn_ref.py reproduces the reference's 1.4 MB
    y-string and its z-string bit for bit (digests recorded from the reference by make_golden.py cnn2_full)."""
    import hashlib
    import json

    g = json.load(open(os.path.join(golden_dir, "cnn2_full.json")))
    sd = weights.seeded_state_dict(cnn_ref.template_state_dict(), seed=0, stress=True)
    x = F.pad(weights.seeded_image((1, 3, 800, 1216), seed=5), (0, 0, 16, 16))
    c = cnn_ref.compress(sd, x)
    assert list(c["shape"]) == g["shape"] == [13, 19]
    assert len(c["strings"][0][0]) == g["y_bytes"] and len(c["strings"][1][0]) == g["z_bytes"]
    assert hashlib.sha1(c["strings"][0][0]).hexdigest() == g["y_sha1"]
    assert hashlib.sha1(c["strings"][1][0]).hexdigest() == g["z_sha1"]


def test_stf_restatement_at_config2_size(golden_dir):
    """One 3x768x512 image (BASELINE.json configs[1]): oracle/stf_ref.py reproduces the reference's y- and z-strings bit
    for bit (digests recorded from the reference by make_golden.py stf_full)."""
    import hashlib
    import json

    g = json.load(open(os.path.join(golden_dir, "stf_full.json")))
    sd = weights.seeded_state_dict(stf_ref.template_state_dict(), seed=0, stress=True)
    c = stf_ref.compress(sd, weights.seeded_image((1, 3, 768, 512), seed=9))
    assert list(c["shape"]) == g["shape"] == [12, 8]
    assert len(c["strings"][0][0]) == g["y_bytes"] and len(c["strings"][1][0]) == g["z_bytes"]
    assert hashlib.sha1(c["strings"][0][0]).hexdigest() == g["y_sha1"]
    assert hashlib.sha1(c["strings"][1][0]).hexdigest() == g["z_sha1"]
