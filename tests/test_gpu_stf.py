"""End-to-end STF (compressai.models.SymmetricalTransFormer on the CUDA kernels) vs fixtures recorded from
the reference model run on the CPU in fp32 (oracle/make_golden.py) and vs the pinned functional oracle.

Tolerances (north_star): x_hat within 0.01 dB PSNR; bit-exact strings where inputs are identical
(self-consistency, batch invariance); likelihoods are checked at 1e-3 on identical inputs in
test_gpu_entropy.py, here only the total rate is compared (the transforms run with bf16 operands)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
PSNR_TOL_DB = 0.01


def psnr(a, b):
    return float(-10 * torch.log10(torch.mean((a.float().cpu() - b.float().cpu()) ** 2)))


@pytest.fixture(scope="module")
def model():
    from compressai.zoo import models
    from oracle import stf_ref, weights

    m = models["stf"]()
    sd = weights.seeded_state_dict(stf_ref.template_state_dict(), seed=0, stress=True)
    missing = m.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys
    m.update(force=True)
    return m.cuda().eval()


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "stf_small.npz"))


@pytest.fixture(scope="module")
def x():
    from oracle import weights

    return weights.seeded_image((1, 3, 128, 192), seed=0)


def test_tables_match_reference_update(model, gold):
    eb = model.entropy_bottleneck
    assert np.array_equal(eb.quantized_cdf.cpu().numpy(), gold["eb_cdf"])
    assert np.array_equal(eb.cdf_length.cpu().numpy(), gold["eb_len"]) and np.array_equal(eb.offset.cpu().numpy(), gold["eb_off"])


def test_analysis_and_hyper_transforms_close_to_reference(model, gold, x):
    B = 1
    y, h, w = model._analysis(x.cuda())
    assert (h, w) == (8, 12)
    y_ref = torch.from_numpy(gold["y"]).permute(0, 2, 3, 1).reshape(-1, 384)
    rel = float((y.cpu() - y_ref).norm() / y_ref.norm())
    assert rel < 2e-2, rel
    z, zh, zw = model._hyper_analysis(y, B, h, w)
    z_ref = torch.from_numpy(gold["z"]).permute(0, 2, 3, 1).reshape(-1, 192)
    rel = float((z.cpu() - z_ref).norm() / z_ref.norm())
    assert (zh, zw) == (2, 3) and rel < 3e-2, rel


def test_forward_matches_reference_quality_and_rate(model, gold, x):
    out = model(x.cuda())
    assert out["x_hat"].shape == (1, 3, 128, 192) and out["likelihoods"]["y"].shape == (1, 384, 8, 12)
    assert out["likelihoods"]["z"].shape == (1, 192, 2, 3)
    x_ref = torch.from_numpy(gold["x_hat"])
    assert abs(psnr(x, out["x_hat"]) - psnr(x, x_ref)) < PSNR_TOL_DB, (psnr(x, out["x_hat"]), psnr(x, x_ref))
    bits = lambda l: float(-torch.log2(torch.as_tensor(l).float().cpu()).sum())
    by, by_ref = bits(out["likelihoods"]["y"]), bits(gold["y_lik"])
    bz, bz_ref = bits(out["likelihoods"]["z"]), bits(gold["z_lik"])
    assert abs(by - by_ref) / by_ref < 2e-2, (by, by_ref)
    assert abs(bz - bz_ref) / bz_ref < 2e-2, (bz, bz_ref)


def test_compress_decompress_self_consistency(model, gold, x):
    """The identity the reference's eval relies on: decompress(compress(x)) == clamp(forward(x).x_hat), exactly."""
    xc = x.cuda()
    c = model.compress(xc)
    assert list(c["shape"]) == [2, 3] and len(c["strings"]) == 2 and len(c["strings"][0]) == 1 and len(c["strings"][1]) == 1
    d = model.decompress(c["strings"], c["shape"])
    f = model(xc)
    assert torch.equal(d["x_hat"], f["x_hat"].clamp(0, 1))
    # stream sizes are in the reference's ballpark (same weights, same image; bf16 transforms)
    ny, nz = len(c["strings"][0][0]), len(c["strings"][1][0])
    assert abs(ny - gold["y_string"].size) / gold["y_string"].size < 3e-2, (ny, gold["y_string"].size)
    assert abs(nz - gold["z_string"].size) / gold["z_string"].size < 5e-2, (nz, gold["z_string"].size)
    assert abs(psnr(x, d["x_hat"]) - psnr(x, torch.from_numpy(gold["x_hat"]).clamp(0, 1))) < PSNR_TOL_DB


def test_batch_strings_equal_single_image_strings(model):
    """Config 3 semantics: per-image strings of a batch == the B=1 strings of each image (bit-exact), and the
    batch decodes back to the per-image reconstructions."""
    from oracle import weights

    xs = torch.cat([weights.seeded_image((1, 3, 128, 128), seed=s) for s in (1, 2, 3)]).cuda()
    cb = model.compress(xs)
    assert len(cb["strings"][0]) == 3 and len(cb["strings"][1]) == 3
    db = model.decompress(cb["strings"], cb["shape"])
    for b in range(3):
        c1 = model.compress(xs[b:b + 1])
        assert c1["strings"][0][0] == cb["strings"][0][b]
        assert c1["strings"][1][0] == cb["strings"][1][b]
        d1 = model.decompress(c1["strings"], c1["shape"])
        assert torch.equal(d1["x_hat"][0], db["x_hat"][b])


def test_micro_batched_streams_give_identical_results(model):
    """Large batches are split over CUDA streams so the coders overlap the transforms; the split must not
    change a single byte, and the device-resident hand-off must decode to the same images."""
    from oracle import weights

    xs = torch.cat([weights.seeded_image((1, 3, 64, 128), seed=10 + s) for s in range(17)]).cuda()
    assert len(model._part_ranges(17)) == 2
    split = model.compress(xs)
    model.micro_batches = 1
    try:
        whole = model.compress(xs)
        d_whole = model.decompress(whole["strings"], whole["shape"])
    finally:
        model.micro_batches = 2
    assert split["strings"] == whole["strings"]
    d_split = model.decompress(split["strings"], split["shape"])
    assert torch.equal(d_split["x_hat"], d_whole["x_hat"])
    dev = model.compress(xs, device_strings=True)
    d_dev = model.decompress(dev["strings"], dev["shape"])
    assert torch.equal(d_dev["x_hat"], d_whole["x_hat"])


def test_grouped_launches_do_not_change_a_byte(model):
    """Grouped conv launches and the one-step coding of the slices >= max_support_slices (icm_conv2d_grouped,
    models/_context.py) against the slice-by-slice loop of stf.py:611-631 / :754-776: strings, x_hat and likelihoods equal."""
    from oracle import weights

    xs = torch.cat([weights.seeded_image((1, 3, 128, 64), seed=30 + s) for s in range(3)]).cuda()
    assert model.grouped and model._tail_grouped()
    a = model.compress(xs)
    da = model.decompress(a["strings"], a["shape"])
    fa = model(xs)
    model.grouped = False
    try:
        b = model.compress(xs)
        db = model.decompress(b["strings"], b["shape"])
        fb = model(xs)
        mixed = model.decompress(a["strings"], a["shape"])  # encoder grouped, decoder not
    finally:
        model.grouped = True
    assert a["strings"] == b["strings"]
    assert torch.equal(da["x_hat"], db["x_hat"]) and torch.equal(mixed["x_hat"], da["x_hat"])
    assert torch.equal(fa["x_hat"], fb["x_hat"])
    assert torch.equal(fa["likelihoods"]["y"], fb["likelihoods"]["y"]) and torch.equal(fa["likelihoods"]["z"], fb["likelihoods"]["z"])


def test_strings_decode_with_the_cpu_oracle(model, x):
    """Cross-implementation check: the GPU-produced y-string of an image decodes, with the pinned CPU coder
    and the GPU-side indexes, to the symbols the GPU encoder consumed."""
    from oracle import coder, entropy

    xc = x.cuda()
    B = 1
    y, h, w = model._analysis(xc)
    z, zh, zw = model._hyper_analysis(y, B, h, w)
    c = model.compress(xc)
    # recompute symbols/indexes exactly as compress does
    from compressai._native import NULL_VIEW, check, lib, stream_ptr, view_bcp

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
    assert int(idx.unique().numel()) > 20  # the stress weights exercise many CDF tables


def test_input_validation(model):
    with pytest.raises(ValueError):
        model.compress(torch.zeros(1, 3, 100, 128, device="cuda"))
    with pytest.raises(ValueError):
        model.decompress([[b"\0" * 8], [b"\0" * 8, b"\0" * 8]], (2, 2))


def test_full_size_batch_strings_equal_single_image_strings(model):
    """BASELINE.json configs[2] semantics at the full 768x512 size: every image of a batch gets the strings it would get
    alone (bit-exact), and the batch decodes to the per-image reconstructions."""
    from oracle import weights

    xs = torch.cat([weights.seeded_image((1, 3, 768, 512), seed=20 + s) for s in range(5)]).cuda()
    cb = model.compress(xs)
    assert [len(s) for s in cb["strings"]] == [5, 5] and list(cb["shape"]) == [12, 8]
    for b in (0, 3):
        c1 = model.compress(xs[b : b + 1])
        assert c1["strings"][0][0] == cb["strings"][0][b] and c1["strings"][1][0] == cb["strings"][1][b]
        d1 = model.decompress(c1["strings"], c1["shape"])
        db = model.decompress([[cb["strings"][0][b]], [cb["strings"][1][b]]], cb["shape"])
        assert torch.equal(d1["x_hat"], db["x_hat"])
    d = model.decompress(cb["strings"], cb["shape"])
    assert torch.equal(d["x_hat"], model(xs)["x_hat"].clamp(0, 1))
    assert 608256 == 384 * 48 * 32 + 192 * 12 * 8  # symbols per image behind the Msym/s figures


@pytest.mark.parametrize("chains,lag", [(0, 0), (2, 2), (1, 3)])
def test_stream_pipeline_equals_plain_api(model, chains, lag):
    """RoundTripPipeline (jobs on a pool of streams, event-chained phases, host I/O) is bit-identical to compress() +
    decompress() on the whole batch: same byte strings, same reconstructions."""
    from compressai.utils.pipeline import RoundTripPipeline
    from oracle import weights

    batches = [torch.cat([weights.seeded_image((1, 3, 64, 128), seed=40 + 7 * k + s) for s in range(6)]) for k in range(3)]
    pipe = RoundTripPipeline(model, n_streams=4, part=4, decoder_streams_per_cta=4, lag=lag, chains=chains)
    dev_batches = [b.cuda() for b in batches]
    x_hats, none = pipe.roundtrip(dev_batches)
    assert none is None and len(x_hats) == 3
    pinned = [b.pin_memory() for b in batches]
    outs = [torch.empty_like(b).pin_memory() for b in batches]
    x_none, strings = pipe.roundtrip(pinned, host_io=True, out_host=outs)
    torch.cuda.synchronize()
    assert x_none is None
    for k, xb in enumerate(dev_batches):
        c = model.compress(xb)
        d = model.decompress(c["strings"], c["shape"])
        assert torch.equal(x_hats[k], d["x_hat"])
        assert torch.equal(outs[k], d["x_hat"].cpu())
        assert strings[k][0] == c["strings"][0] and strings[k][1] == c["strings"][1]


@pytest.mark.parametrize("priority", [False, True])
def test_stream_pipeline_with_cuda_graphs_equals_plain_api(model, priority):
    """RoundTripPipeline(cuda_graphs=True): the first job of a (slot, shape) runs eagerly, the second is captured as four
    CUDA graphs and from then on replayed; every pass -- eager, capturing, replaying; device-resident and host I/O -- must give
    the plain API's strings and reconstructions, for different images each time (static buffers are refilled, not reused)."""
    from compressai.utils.pipeline import RoundTripPipeline
    from oracle import weights

    pipe = RoundTripPipeline(model, n_streams=3, part=4, decoder_streams_per_cta=4, lag=2, chains=2, decode_priority=priority, cuda_graphs=True)
    for rnd in range(4):
        batches = [torch.cat([weights.seeded_image((1, 3, 64, 128), seed=500 + 100 * rnd + 7 * k + s) for s in range(8)]) for k in range(3)]
        dev_batches = [b.cuda() for b in batches]
        host = rnd % 2 == 1
        if host:
            outs = [torch.empty_like(b).pin_memory() for b in batches]
            _, strings = pipe.roundtrip([b.pin_memory() for b in batches], host_io=True, out_host=outs)
            torch.cuda.synchronize()
        else:
            x_hats, _ = pipe.roundtrip(dev_batches)
        for k, xb in enumerate(dev_batches):
            c = model.compress(xb)
            d = model.decompress(c["strings"], c["shape"])
            if host:
                assert torch.equal(outs[k], d["x_hat"].cpu()), (rnd, k)
                assert strings[k][0] == c["strings"][0] and strings[k][1] == c["strings"][1], (rnd, k)
            else:
                assert torch.equal(x_hats[k], d["x_hat"]), (rnd, k)
    assert len(pipe._job_graphs) >= 3  # graphs were captured and used


def test_full_size_stream_sizes_track_the_reference(model, golden_dir):
    """The 768x512 image whose reference strings are recorded in tests/golden/stf_full.json: same latent shape, stream sizes
    within a few percent (bf16 transforms), decompress(compress(x)) == clamp(forward(x))."""
    import json

    from oracle import weights

    g = json.load(open(os.path.join(golden_dir, "stf_full.json")))
    x = weights.seeded_image((1, 3, 768, 512), seed=9).cuda()
    c = model.compress(x)
    assert list(c["shape"]) == g["shape"]
    ny, nz = len(c["strings"][0][0]), len(c["strings"][1][0])
    assert abs(ny - g["y_bytes"]) / g["y_bytes"] < 3e-2, (ny, g["y_bytes"])
    assert abs(nz - g["z_bytes"]) / g["z_bytes"] < 5e-2, (nz, g["z_bytes"])
    d = model.decompress(c["strings"], c["shape"])
    assert torch.equal(d["x_hat"], model(x)["x_hat"].clamp(0, 1))
