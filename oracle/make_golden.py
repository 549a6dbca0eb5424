"""Generate tests/golden/* from the REFERENCE ITSELF (run in the build container only).

    python -m oracle.make_golden          # from the repo root; needs /root/reference
    python -m oracle.make_golden cnn      # only tests/golden/cnn_small.npz (WACNN, 256x256 = BASELINE configs[0])
    python -m oracle.make_golden stf_full   # only tests/golden/stf_full.json (768x512 = BASELINE configs[1], digests)
    python -m oracle.make_golden cnn2_full  # only tests/golden/cnn2_full.json (832x1216 = BASELINE configs[3], digests)

Sources of truth used here:
  * the reference's shipped native binaries, executed through oracle/refbin.py (rANS, pmf->cdf);
  * the reference's unmodified Python modules, imported through oracle/refshim.py
    (entropy_models.py, models/stf.py), on the CPU in fp32.
Nothing under oracle/*.c / oracle/entropy.py / oracle/stf_ref.py (the restatements) is used to
PRODUCE a fixture; tests/test_oracle_pinned.py replays the fixtures against those restatements.
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)
GOLD = os.path.join(REPO, "tests", "golden")

from oracle import refbin, refshim, weights  # noqa: E402


def sha1(b):
    return hashlib.sha1(b).hexdigest()


def seeded_stream(n, seed, table, kind="uniform"):
    """SURVEY.md §8c/§8d synthetic (symbols, indexes)."""
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        idx = rng.integers(0, 64, n)
    elif kind == "lowrate":
        idx = np.minimum(rng.geometric(0.15, n) - 1, 63)
    elif kind == "zero":
        idx = np.zeros(n, np.int64)
    sym = np.rint(rng.normal(0, table[idx])).astype(np.int32)
    if kind == "zero":
        sym[:] = 0
    return sym, idx.astype(np.int32)


def main():
    os.makedirs(GOLD, exist_ok=True)
    h = refshim.install("binary")
    stf, em = h["stf"], h["entropy_models"]
    torch.manual_seed(0)

    # ---------------------------------------------------------------- tables (reference update())
    gc = em.GaussianConditional(None)
    gc.update_scale_table(stf.get_scale_table())
    cdf = gc._quantized_cdf.numpy().astype(np.int32)
    lengths = gc._cdf_length.numpy().astype(np.int32)
    offsets = gc._offset.numpy().astype(np.int32)
    table = gc.scale_table.numpy()
    assert sha1(cdf.astype("<i4").tobytes()) == "5000194aa652d73520578889b0d0b77c61a7d389"

    kat = {"pmf": [], "rans_small": [], "streams": [], "gc_table_sha1": sha1(cdf.astype("<i4").tobytes())}
    # ---------------------------------------------------------------- R5 KATs
    rng = np.random.default_rng(7)
    pmfs = [[0.25, 0.5, 0.25], [0.1, 0.2, 0.3, 0.4], [1e-9, 0.999999, 1e-9, 1e-9], [0.5, 0, 0, 0.5], [1 / 3, 1 / 3, 1 / 3]]
    for n in (2, 7, 33, 257):
        p = rng.random(n) ** 4
        p[rng.integers(0, n, max(1, n // 5))] = 0.0
        pmfs.append((p / p.sum()).astype(np.float32).tolist())
    for p in pmfs:
        for prec in (16, 12):
            if len(p) + 1 > (1 << prec):
                continue
            kat["pmf"].append({"pmf": [float(np.float32(v)) for v in p], "precision": prec,
                               "cdf": refbin.pmf_to_quantized_cdf(p, prec).tolist()})
    # ---------------------------------------------------------------- R1-R4 small KATs
    small_cdfs = np.array([[0, 16384, 49152, 65536, 0, 0], [0, 1, 32768, 65535, 65536, 0]], np.int32)
    small = [([-1, 0, -1, 0], [0, 0, 0, 0]), ([0, 1, -1, 0, 1, -1, 0, 0], [1, 1, 1, 0, 1, 0, 1, 0]),
             ([1, 0, 0, 0], [0, 0, 0, 0]), ([-2, 0, 0, 0], [0, 0, 0, 0]), ([100000, 0, 0, 0], [0, 0, 0, 0]),
             ([-100000, 0, 0, 0], [1, 0, 0, 0]), ([0, 0, 0], [0, 1, 0]), ([7, -9, 33, -1000, 12345678, 0, 2], [1, 0, 1, 0, 1, 0, 1]),
             ([15, -15, 16, -16, 255, -256, 4095, 4096, 65535, -65536, 1 << 20, -(1 << 20), (1 << 26), -(1 << 26)], [0] * 14)]
    for s, i in small:
        b = refbin.RansEncoder().encode_with_indexes(s, i, small_cdfs, [4, 5], [-1, -1])
        d = refbin.RansDecoder().decode_with_indexes(b, i, small_cdfs, [4, 5], [-1, -1])
        assert d.tolist() == list(s)
        kat["rans_small"].append({"symbols": s, "indexes": i, "hex": b.hex()})
    kat["small_tables"] = {"cdfs": small_cdfs.tolist(), "sizes": [4, 5], "offsets": [-1, -1]}
    # ---------------------------------------------------------------- seeded streams on the real GC tables
    for n, seed, kind in [(1000, 1, "uniform"), (49152, 2, "uniform"), (589824, 1234, "uniform"), (49152, 3, "lowrate"),
                          (49152, 4, "zero"), (4096, 5, "uniform")]:
        sym, idx = seeded_stream(n, seed, table, kind)
        if seed == 5:  # adversarial: bypass boundaries and long escapes
            c = -offsets[idx]
            sym = np.where(np.arange(n) % 4 == 0, c, np.where(np.arange(n) % 4 == 1, -c, np.where(np.arange(n) % 4 == 2, c + 1, -c - 1))).astype(np.int32)
            sym[::97] = 100000
            sym[1::97] = -100000
        b = refbin.RansEncoder().encode_with_indexes(sym, idx, cdf, lengths, offsets)
        d = refbin.RansDecoder().decode_with_indexes(b, idx, cdf, lengths, offsets)
        assert np.array_equal(d, sym)
        # split decode (decode_stream x3) and buffered multi-call encode behave like one-shot
        e = refbin.BufferedRansEncoder()
        k = n // 3
        e.encode_with_indexes(sym[:k], idx[:k], cdf, lengths, offsets)
        e.encode_with_indexes(sym[k:], idx[k:], cdf, lengths, offsets)
        assert e.flush() == b
        kat["streams"].append({"n": n, "seed": seed, "kind": kind if seed != 5 else "adversarial",
                               "nbytes": len(b), "sha1": sha1(b), "sym_sha1": sha1(sym.astype("<i4").tobytes())})
    with open(os.path.join(GOLD, "rans_kat.json"), "w") as f:
        json.dump(kat, f)

    # pmf rows (floats) for three tables, so pmf->cdf can be replayed without torch's erfc
    mult = -gc._standardized_quantile(gc.tail_mass / 2)
    np.savez_compressed(
        os.path.join(GOLD, "gc_tables.npz"), scale_table=table, cdf_length=lengths, offset=offsets,
        row0=cdf[0, : lengths[0]], row10=cdf[10, : lengths[10]], row40=cdf[40, : lengths[40]], row63=cdf[63, : lengths[63]],
        multiplier=np.float64(mult),
    )

    # ---------------------------------------------------------------- entropy-model KATs (reference python)
    g = torch.Generator().manual_seed(11)
    logu = torch.exp(torch.empty(20000).uniform_(np.log(0.01), np.log(400.0), generator=g))
    t = torch.from_numpy(table)
    edge = torch.cat([t, torch.nextafter(t, torch.tensor(0.0)), torch.nextafter(t, torch.tensor(1e9)),
                      torch.tensor([0.0, -1.0, 0.11, 256.0, 1e9, 0.10999999])])
    scales = torch.cat([logu, edge])
    idx_ref = gc.build_indexes(scales).numpy().astype(np.int32)
    means = torch.randn(scales.shape, generator=g) * 2
    y = means + torch.randn(scales.shape, generator=g) * scales.clamp(0.05, 50)
    y[:6] = torch.tensor([0.5, 1.5, 2.5, -0.5, -1.5, -2.5]) + means[:6]
    gc.eval()
    with torch.no_grad():
        y_hat, lik = gc(y, scales, means)
        sym = gc.quantize(y, "symbols", means)
    eb = em.EntropyBottleneck(192)
    ebsd = weights.seeded_state_dict(eb.state_dict(), seed=3, stress=False)
    eb.load_state_dict({k: v for k, v in ebsd.items() if k in dict(eb.named_parameters())}, strict=False)
    eb.update(force=True)
    eb.eval()
    z = torch.randn(2, 192, 3, 5, generator=g) * 4
    with torch.no_grad():
        z_hat, z_lik = eb(z)
        z_strings = eb.compress(z)
        z_dec = eb.decompress(z_strings, z.shape[-2:])
    assert torch.equal(z_dec, z_hat)
    np.savez_compressed(
        os.path.join(GOLD, "entropy_kat.npz"),
        scales=scales.numpy(), indexes=idx_ref, means=means.numpy(), y=y.numpy(), y_hat=y_hat.numpy(),
        y_lik=lik.numpy(), symbols=sym.numpy(),
        eb_cdf=eb._quantized_cdf.numpy(), eb_len=eb._cdf_length.numpy(), eb_off=eb._offset.numpy(),
        z=z.numpy(), z_hat=z_hat.numpy(), z_lik=z_lik.numpy(),
        z_string0=np.frombuffer(z_strings[0], np.uint8), z_string1=np.frombuffer(z_strings[1], np.uint8),
    )

    # ---------------------------------------------------------------- model-level: STF 128x192, seeded stress weights
    m = stf.SymmetricalTransFormer().eval()
    sd = weights.seeded_state_dict(m.state_dict(), seed=0, stress=True)
    m.load_state_dict(sd)
    m.update(force=True)
    x = weights.seeded_image((1, 3, 128, 192), seed=0)
    rec = {}

    def hook(name):
        def f(mod, inp, out):
            rec.setdefault(name, []).append(out.detach().clone())
        return f

    hs = []
    for i in range(12):
        hs.append(m.cc_mean_transforms[i].register_forward_hook(hook("mu")))
        hs.append(m.cc_scale_transforms[i].register_forward_hook(hook("scale")))
    hs.append(m.h_a.register_forward_hook(hook("z")))
    hs.append(m.h_mean_s.register_forward_hook(hook("lm")))
    hs.append(m.h_scale_s.register_forward_hook(hook("ls")))
    hs.append(m.layers[3].register_forward_hook(lambda mod, i, o: rec.setdefault("y_tok", []).append(o[0].detach().clone())))
    with torch.no_grad():
        out = m(x)
        fwd = {k: [t.clone() for t in v] for k, v in rec.items()}
        rec.clear()
        c = m.compress(x)
        d = m.decompress(c["strings"], c["shape"])
    for hh in hs:
        hh.remove()
    assert torch.equal(d["x_hat"], out["x_hat"].clamp(0, 1))
    y_tok = fwd["y_tok"][0]
    yv = y_tok.view(1, 8, 12, 384).permute(0, 3, 1, 2).contiguous()
    mu = torch.cat(fwd["mu"], 1)
    sc = torch.cat(fwd["scale"], 1)
    sym = torch.round(yv - mu).int()
    idx = m.gaussian_conditional.build_indexes(sc)
    print("stf_small: y-string", len(c["strings"][0][0]), "B z-string", len(c["strings"][1][0]), "B; distinct idx", idx.unique().numel(),
          "max|sym|", int(sym.abs().max()), "psnr", float(-10 * torch.log10(torch.mean((x - d['x_hat']) ** 2))))
    np.savez_compressed(
        os.path.join(GOLD, "stf_small.npz"),
        y=yv.numpy(), z=fwd["z"][0].numpy(), latent_means=fwd["lm"][0].numpy(), latent_scales=fwd["ls"][0].numpy(),
        mu=mu.numpy(), scale=sc.numpy(), symbols=sym.numpy().astype(np.int32), indexes=idx.numpy().astype(np.uint8),
        x_hat=out["x_hat"].numpy(), y_lik=out["likelihoods"]["y"].numpy(), z_lik=out["likelihoods"]["z"].numpy(),
        y_string=np.frombuffer(c["strings"][0][0], np.uint8), z_string=np.frombuffer(c["strings"][1][0], np.uint8),
        eb_cdf=m.entropy_bottleneck._quantized_cdf.numpy(), eb_len=m.entropy_bottleneck._cdf_length.numpy(),
        eb_off=m.entropy_bottleneck._offset.numpy(),
    )
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


def cnn_golden():
    """Model-level fixture for WACNN (cnn.py) at 1x3x256x256, seeded stress weights."""
    h = refshim.install("binary")
    cnn = h["cnn"]
    torch.manual_seed(0)
    m = cnn.WACNN().eval()
    sd = weights.seeded_state_dict(m.state_dict(), seed=0, stress=True)
    m.load_state_dict(sd)
    m.update(force=True)
    x = weights.seeded_image((1, 3, 256, 256), seed=0)
    rec = {}

    def hook(name):
        def f(mod, inp, out):
            rec.setdefault(name, []).append(out.detach().clone())
        return f

    hs = []
    for i in range(10):
        hs.append(m.cc_mean_transforms[i].register_forward_hook(hook("mu")))
        hs.append(m.cc_scale_transforms[i].register_forward_hook(hook("scale")))
    for name in ("g_a", "h_a", "h_mean_s", "h_scale_s"):
        hs.append(getattr(m, name).register_forward_hook(hook(name)))
    hs.append(m.g_a[1].register_forward_hook(hook("gdn0")))
    hs.append(m.g_a[4].register_forward_hook(hook("win0")))
    with torch.no_grad():
        out = m(x)
        fwd = {k: [t.clone() for t in v] for k, v in rec.items()}
        rec.clear()
        c = m.compress(x)
        d = m.decompress(c["strings"], c["shape"])
    for hh in hs:
        hh.remove()
    assert torch.equal(d["x_hat"], out["x_hat"].clamp(0, 1))
    y = fwd["g_a"][0]
    mu = torch.cat(fwd["mu"], 1)
    sc = torch.cat(fwd["scale"], 1)
    sym = torch.round(y - mu).int()
    idx = m.gaussian_conditional.build_indexes(sc)
    print("cnn_small: y-string", len(c["strings"][0][0]), "B z-string", len(c["strings"][1][0]), "B; distinct idx", idx.unique().numel(),
          "max|sym|", int(sym.abs().max()), "psnr", float(-10 * torch.log10(torch.mean((x - d['x_hat']) ** 2))))
    np.savez_compressed(
        os.path.join(GOLD, "cnn_small.npz"),
        y=y.numpy(), z=fwd["h_a"][0].numpy(), latent_means=fwd["h_mean_s"][0].numpy(), latent_scales=fwd["h_scale_s"][0].numpy(),
        gdn0=fwd["gdn0"][0][:, :, ::4, ::4].numpy(), win0=fwd["win0"][0][:, :, ::2, ::2].numpy(),
        mu=mu.numpy(), scale=sc.numpy(), symbols=sym.numpy().astype(np.int32), indexes=idx.numpy().astype(np.uint8),
        x_hat=out["x_hat"].numpy(), y_lik=out["likelihoods"]["y"].numpy(), z_lik=out["likelihoods"]["z"].numpy(),
        y_string=np.frombuffer(c["strings"][0][0], np.uint8), z_string=np.frombuffer(c["strings"][1][0], np.uint8),
        eb_cdf=m.entropy_bottleneck._quantized_cdf.numpy(), eb_len=m.entropy_bottleneck._cdf_length.numpy(),
        eb_off=m.entropy_bottleneck._offset.numpy(),
    )
    print("cnn_small.npz", os.path.getsize(os.path.join(GOLD, "cnn_small.npz")))


def cnn2_full_golden():
    """BASELINE.json configs[3] size: the WACNN2 codec (identical layers to WACNN) on a 800x1216 image zero-padded to
    832x1216, compressed by the reference on the CPU.  Only digests are stored (the y-string is 1.4 MB)."""
    import torch.nn.functional as F

    h = refshim.install("binary")
    torch.manual_seed(0)
    m = h["cnn"].WACNN().eval()
    sd = weights.seeded_state_dict(m.state_dict(), seed=0, stress=True)
    m.load_state_dict(sd)
    m.update(force=True)
    x = F.pad(weights.seeded_image((1, 3, 800, 1216), seed=5), (0, 0, 16, 16))
    with torch.no_grad():
        c = m.compress(x)
    out = {"image": "seeded_image((1,3,800,1216), seed=5) padded (0,0,16,16)", "shape": list(c["shape"]),
           "y_bytes": len(c["strings"][0][0]), "z_bytes": len(c["strings"][1][0]),
           "y_sha1": sha1(c["strings"][0][0]), "z_sha1": sha1(c["strings"][1][0])}
    with open(os.path.join(GOLD, "cnn2_full.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("cnn2_full:", out)


def stf_full_golden():
    """BASELINE.json configs[1] size: STF on one 3x768x512 image, compressed by the reference on the CPU (digests only)."""
    h = refshim.install("binary")
    torch.manual_seed(0)
    m = h["stf"].SymmetricalTransFormer().eval()
    sd = weights.seeded_state_dict(m.state_dict(), seed=0, stress=True)
    m.load_state_dict(sd)
    m.update(force=True)
    x = weights.seeded_image((1, 3, 768, 512), seed=9)
    with torch.no_grad():
        c = m.compress(x)
    out = {"image": "seeded_image((1,3,768,512), seed=9)", "shape": list(c["shape"]),
           "y_bytes": len(c["strings"][0][0]), "z_bytes": len(c["strings"][1][0]),
           "y_sha1": sha1(c["strings"][0][0]), "z_sha1": sha1(c["strings"][1][0])}
    with open(os.path.join(GOLD, "stf_full.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("stf_full:", out)


def eval_golden():
    """The reference's OWN `inference()` and `psnr()` (compressai/utils/eval_model/__main__.py:78-81, 97-140), executed
    from their unmodified source text on the reference STF: pad-to-64 / compress / decompress / crop / psnr / bpp of an
    image whose size is not a multiple of 64.  The module itself cannot be imported here (pytorch_msssim, pycocotools,
    detectron2, ptflops ... at its top), so the two function definitions are cut out of the file with `ast` and
    exec'ed; `reconstruct` (a PNG write through torchvision) is replaced by a no-op."""
    import ast
    import math
    import time

    import torch.nn.functional as F

    path = os.path.join(refshim.REF, "compressai", "utils", "eval_model", "__main__.py")
    src = open(path).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "F": F, "math": math, "os": os, "time": time, "reconstruct": lambda *a, **k: None}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("psnr", "inference"):
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    h = refshim.install("binary")
    torch.manual_seed(0)
    m = h["stf"].SymmetricalTransFormer().eval()
    m.load_state_dict(weights.seeded_state_dict(m.state_dict(), seed=0, stress=True))
    m.update(force=True)
    x = weights.seeded_image((3, 100, 150), seed=21)
    captured = {}
    orig_c, orig_d = m.compress, m.decompress

    def cap_c(xp):
        captured["x_padded_shape"] = list(xp.shape)
        captured["enc"] = orig_c(xp)
        return captured["enc"]

    def cap_d(strings, shape):
        captured["dec"] = orig_d(strings, shape)
        captured["x_hat_padded"] = captured["dec"]["x_hat"].clone()
        return captured["dec"]

    m.compress, m.decompress = cap_c, cap_d
    import tempfile

    rv = ns["inference"](m, x, "img.png", tempfile.mkdtemp(prefix="recon_"))
    enc = captured["enc"]
    np.savez_compressed(
        os.path.join(GOLD, "eval_small.npz"),
        psnr=np.float64(rv["psnr"]), bpp=np.float64(rv["bpp"]), x_padded_shape=np.asarray(captured["x_padded_shape"]),
        shape=np.asarray(list(enc["shape"])), y_string=np.frombuffer(enc["strings"][0][0], np.uint8),
        z_string=np.frombuffer(enc["strings"][1][0], np.uint8), x_hat_padded=captured["x_hat_padded"].numpy().astype(np.float32),
        x_hat=captured["dec"]["x_hat"].numpy().astype(np.float32))
    print("eval_small: psnr %.4f bpp %.4f padded %s y %d B z %d B" % (rv["psnr"], rv["bpp"], captured["x_padded_shape"],
                                                                    len(enc["strings"][0][0]), len(enc["strings"][1][0])))
    refshim.uninstall(h)



TRAIN_PROBES = ["h_a.8.bias", "cc_scale_transforms.3.8.bias", "cc_mean_transforms.0.0.bias", "lrp_transforms.11.8.bias",
                "patch_embed.proj.bias", "layers.0.blocks.1.attn.relative_position_bias_table", "syn_layers.3.blocks.1.mlp.fc2.bias",
                "layers.2.blocks.3.norm1.weight", "entropy_bottleneck._bias0", "entropy_bottleneck._matrix2", "entropy_bottleneck.quantiles",
                "end_conv.2.bias"]


def train_golden():
    """One rate-distortion training step of the reference STF on the CPU (BASELINE.json configs[4] at a small size):
    the UNMODIFIED reference modules in .train() (DropPath 0.2, noise quantisation; stf.py:582-645), the loss of
    train.py:53-74 read with the model's "x_hat" key, clip_grad_norm_ 1.0, Adam 1e-5 on everything but the quantiles and
    Adam 1e-4 on the quantiles after aux_loss.backward() (train.py:161-168, 196-214).  Records the loss terms, the total
    gradient norm, the gradients of a few small parameters and those parameters after the step."""
    import math

    h = refshim.install("binary")
    torch.manual_seed(0)
    m = h["stf"].SymmetricalTransFormer()
    m.load_state_dict(weights.seeded_state_dict(m.state_dict(), seed=0, stress=True))
    m.train()
    x = weights.seeded_image((2, 3, 128, 128), seed=41)
    named = dict(m.named_parameters())
    main = [p for n, p in named.items() if not n.endswith(".quantiles")]
    aux = [p for n, p in named.items() if n.endswith(".quantiles")]
    opt, aux_opt = torch.optim.Adam(main, lr=1e-5), torch.optim.Adam(aux, lr=1e-4)
    lmbda = 800.0
    out = {}
    for step in range(2):
        torch.manual_seed(4242 + step)
        opt.zero_grad()
        aux_opt.zero_grad()
        o = m(x)
        N, _, H, W = x.shape
        bpp = sum(torch.log(l).sum() / (-math.log(2) * N * H * W) for l in o["likelihoods"].values())
        mse = torch.nn.functional.mse_loss(x, o["x_hat"])
        loss = lmbda * mse + bpp
        loss.backward()
        if step == 0:
            for n in TRAIN_PROBES:
                out["grad/" + n] = named[n].grad.detach().clone().numpy()
                out["before/" + n] = named[n].detach().clone().numpy()
        norm = torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
        opt.step()
        a = m.aux_loss()
        a.backward()
        if step == 0:
            out["aux_grad/entropy_bottleneck.quantiles"] = named["entropy_bottleneck.quantiles"].grad.detach().clone().numpy()
        aux_opt.step()
        out[f"loss{step}"], out[f"bpp{step}"], out[f"mse{step}"] = np.float64(loss.item()), np.float64(bpp.item()), np.float64(mse.item())
        out[f"aux{step}"], out[f"norm{step}"] = np.float64(a.item()), np.float64(float(norm))
        out[f"y_lik_logsum{step}"] = np.float64(torch.log(o["likelihoods"]["y"]).sum().item())
        out[f"z_lik_logsum{step}"] = np.float64(torch.log(o["likelihoods"]["z"]).sum().item())
        if step == 0:
            for n in TRAIN_PROBES:
                out["after/" + n] = named[n].detach().clone().numpy()
        print(f"train step {step}: loss {loss.item():.6f} bpp {bpp.item():.6f} mse {mse.item():.6f} aux {a.item():.4f} |g| {float(norm):.4f}")
    np.savez_compressed(os.path.join(GOLD, "train_step.npz"), **out)
    refshim.uninstall(h)


if __name__ == "__main__":
    if sys.argv[1:] == ["train"]:
        train_golden()
    elif sys.argv[1:] == ["eval"]:
        eval_golden()
    elif sys.argv[1:] == ["stf_full"]:
        stf_full_golden()
    elif sys.argv[1:] == ["cnn"]:
        cnn_golden()
    elif sys.argv[1:] == ["cnn2_full"]:
        cnn2_full_golden()
    else:
        main()
        cnn_golden()
        cnn2_full_golden()
        stf_full_golden()
        eval_golden()
        train_golden()
