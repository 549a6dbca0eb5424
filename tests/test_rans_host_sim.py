"""Host simulation of the bucket-table rANS decoder (csrc/rans_lane.cuh: the per-symbol code that
rans_decode_bucket_kernel runs) against the pinned CPU oracle and the reference-binary fixtures -- bit-exact, no GPU.

Covers what the GPU tests cannot easily isolate: the bucket-table image (bucket bits chosen under a shared-memory
budget), the in-register resolution of up to three symbols per bucket, the binary-search path for crowded buckets
(forced by a tiny budget), escapes incl. long payloads that drain the stream-word ring, and multi-step decoding with
carried state."""
import ctypes as C
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
BUDGET = 226 * 1024 - 4096 - 8 * 2528  # what icm_tables_create gives the image (csrc/rans.cu)


@pytest.fixture(scope="module")
def sim():
    out_dir = os.path.join(HERE, "host_sim", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "librans_host_sim.so")
    src = os.path.join(HERE, "host_sim", "rans_host_sim.cpp")
    hdr = os.path.join(REPO, "image-compression-for-machine_b200", "csrc", "rans_lane.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "-I", os.path.dirname(hdr), src, "-o", so])
    return C.CDLL(so)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _tables(cdf, lengths, offsets):
    return np.ascontiguousarray(cdf, np.int32), np.ascontiguousarray(lengths, np.int32), np.ascontiguousarray(offsets, np.int32)


def sim_decode(L, tabs, data, idx, steps=None, budget=BUDGET):
    cdf, lengths, offsets = tabs
    idx = np.ascontiguousarray(idx, np.int32)
    words = np.frombuffer(data, np.uint32).copy()
    steps = np.asarray([idx.size] if steps is None else steps, np.int64)
    out = np.zeros(idx.size, np.int32)
    rare = C.c_longlong(0)
    rc = L.sim_decode(_p(cdf), cdf.shape[0], cdf.shape[1], _p(lengths), _p(offsets), C.c_longlong(budget), _p(words), words.size,
                      _p(idx), C.c_longlong(idx.size), _p(steps), steps.size, _p(out), C.byref(rare))
    assert rc >= 0
    sim_decode.rare = rare.value
    return out


@pytest.fixture(scope="module")
def env(golden_dir):
    from oracle import coder, entropy

    kat = json.load(open(os.path.join(golden_dir, "rans_kat.json")))
    cdf, lengths, offsets = entropy.gc_tables()
    return dict(coder=coder, kat=kat, gc=_tables(cdf, lengths, offsets), table=entropy.scale_table().numpy())


def test_small_kats_from_the_reference_binary(sim, env):
    t = env["kat"]["small_tables"]
    width = max(len(r) for r in t["cdfs"])
    cdf = np.zeros((len(t["cdfs"]), width), np.int32)
    for i, r in enumerate(t["cdfs"]):
        cdf[i, :len(r)] = r
    tabs = _tables(cdf, t["sizes"], t["offsets"])
    for k in env["kat"]["rans_small"]:
        assert sim_decode(sim, tabs, bytes.fromhex(k["hex"]), k["indexes"]).tolist() == k["symbols"]


def _stream(k, env):
    from oracle.make_golden import seeded_stream

    cdf, lengths, offsets = env["gc"]
    kind = "uniform" if k["kind"] == "adversarial" else k["kind"]
    sym, idx = seeded_stream(k["n"], k["seed"], env["table"], kind)
    if k["kind"] == "adversarial":
        c = -offsets[idx]
        a = np.arange(k["n"]) % 4
        sym = np.where(a == 0, c, np.where(a == 1, -c, np.where(a == 2, c + 1, -c - 1))).astype(np.int32)
        sym[::97] = 100000
        sym[1::97] = -100000
    return sym, idx


def test_seeded_streams_match_reference_binary(sim, env):
    for k in env["kat"]["streams"]:
        sym, idx = _stream(k, env)
        b = env["coder"].rans_encode(sym, idx, *env["gc"])
        assert len(b) == k["nbytes"] and hashlib.sha1(b).hexdigest() == k["sha1"], k  # the stream the reference binary produced
        n = k["n"]
        got = sim_decode(sim, env["gc"], b, idx, steps=[n // 5, n // 2, n])
        assert np.array_equal(got, sym), k


@pytest.mark.parametrize("budget", [BUDGET, 100 * 1024, 58 * 1024])
def test_every_table_and_every_cum_value(sim, env, budget):
    """For every Gaussian table: symbols chosen so that the decoder sees each bucket and both edges of every
    symbol; small budgets force one-bucket tables, i.e. the binary-search path for nearly every symbol."""
    cdf, lengths, offsets = env["gc"]
    coder = env["coder"]
    rng = np.random.default_rng(11)
    bits = np.zeros(64, np.int32)
    nbytes = sim.sim_image_info(_p(cdf), 64, cdf.shape[1], _p(lengths), _p(offsets), C.c_longlong(budget), _p(bits))
    assert 0 < nbytes <= budget
    assert len(set(bits.tolist())) == 1 and bits[0] == {BUDGET: 7, 100 * 1024: 5, 58 * 1024: 1}[budget]  # 128 / 32 / 2 buckets per table
    for t in range(64):
        nsym = int(lengths[t]) - 1
        v = np.concatenate([np.arange(nsym - 1), rng.integers(0, nsym - 1, 400), [nsym - 1 + 3, -5 + 0]])  # all in-table symbols + escapes
        sym = (v + offsets[t]).astype(np.int32)
        sym[-1] = offsets[t] - 5
        rng.shuffle(sym)
        idx = np.full(sym.size, t, np.int32)
        b = coder.rans_encode(sym, idx, cdf, lengths, offsets)
        assert np.array_equal(sim_decode(sim, env["gc"], b, idx, budget=budget), sym), t


def test_mixed_tables_low_rate_and_mismatched_statistics(sim, env):
    cdf, lengths, offsets = env["gc"]
    coder = env["coder"]
    rng = np.random.default_rng(3)
    n = 20000
    for kind in range(3):
        if kind == 0:    # default-weights-like: table 0, |sym| <= 3 (cum lands on 0 / 65534 / 65535)
            idx = np.zeros(n, np.int32); sym = np.rint(rng.normal(0, 0.9, n)).astype(np.int32)
        elif kind == 1:  # statistics 3x wider than the tables: many escapes, long payloads
            idx = rng.integers(0, 64, n).astype(np.int32); sym = np.rint(rng.normal(0, 3 * env["table"][idx])).astype(np.int32)
        else:
            idx = np.minimum(rng.geometric(0.15, n) - 1, 63).astype(np.int32); sym = np.rint(rng.normal(0, env["table"][idx])).astype(np.int32)
        b = coder.rans_encode(sym, idx, cdf, lengths, offsets)
        assert np.array_equal(sim_decode(sim, env["gc"], b, idx, steps=[1, 2, 7, n // 3, n]), sym)
        if kind == 2:  # data that follows the tables: all but a few per cent resolve in registers from one load
            assert sim_decode.rare < 0.03 * n, sim_decode.rare


def test_entropy_bottleneck_shaped_tables(sim, env):
    """192 tables of 23 entries (the z path): the image must fit and round-trip against the oracle."""
    rng = np.random.default_rng(5)
    n_cdf, size = 192, 23
    cdf = np.zeros((n_cdf, size), np.int32)
    for t in range(n_cdf):
        p = rng.dirichlet(np.full(size - 1, 0.6)) * (65536 - (size - 1))
        f = np.floor(p).astype(np.int64) + 1
        f[np.argmax(f)] += 65536 - f.sum()
        cdf[t, 1:] = np.cumsum(f)
    lengths, offsets = np.full(n_cdf, size, np.int32), np.full(n_cdf, -10, np.int32)
    tabs = _tables(cdf, lengths, offsets)
    idx = np.repeat(np.arange(n_cdf), 40).astype(np.int32)
    sym = rng.integers(-13, 14, idx.size).astype(np.int32)
    b = env["coder"].rans_encode(sym, idx, cdf, lengths, offsets)
    assert np.array_equal(sim_decode(sim, tabs, b, idx), sym)


def test_long_escapes_drain_the_word_ring(sim, env):
    """Every symbol an 8-nibble escape (52 bits each): a 32-symbol chunk consumes more words than the ring guarantees
    ahead, so the out-of-line path has to refill by itself."""
    cdf, lengths, offsets = env["gc"]
    rng = np.random.default_rng(8)
    n = 3000
    idx = rng.integers(0, 64, n).astype(np.int32)
    sym = (rng.integers(2 ** 25, 2 ** 27, n) * rng.choice([-1, 1], n)).astype(np.int32)
    b = env["coder"].rans_encode(sym, idx, cdf, lengths, offsets)
    assert len(b) > 5 * n  # ~46 bits per symbol: 32 symbols consume ~46 words, more than the ring guarantees ahead
    assert np.array_equal(sim_decode(sim, env["gc"], b, idx, steps=[5, 100, n]), sym)
