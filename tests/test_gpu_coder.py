"""GPU rANS coder (csrc/rans.cu through compressai.ans) vs the pinned CPU oracle and the fixtures recorded
from the reference's own binaries.  Bit-exact: every comparison is ==."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env(golden_dir):
    from compressai import ans
    from oracle import coder, entropy

    kat = json.load(open(os.path.join(golden_dir, "rans_kat.json")))
    cdf, lengths, offsets = entropy.gc_tables()
    return dict(ans=ans, coder=coder, kat=kat, gc=(cdf, lengths, offsets), table=entropy.scale_table().numpy(),
                tables=ans.Tables(cdf, lengths, offsets))


def test_reference_call_shapes_with_python_lists(env):
    """Exactly the reference's call style: Python lists in, bytes / list[int] out (SURVEY.md §8c KATs)."""
    ans, kat = env["ans"], env["kat"]
    t = kat["small_tables"]
    for k in kat["rans_small"]:
        b = ans.RansEncoder().encode_with_indexes(k["symbols"], k["indexes"], t["cdfs"], t["sizes"], t["offsets"])
        assert isinstance(b, bytes) and b.hex() == k["hex"]
        d = ans.RansDecoder().decode_with_indexes(b, k["indexes"], t["cdfs"], t["sizes"], t["offsets"])
        assert isinstance(d, list) and d == k["symbols"]
    e = ans.BufferedRansEncoder()
    k = kat["rans_small"][1]
    e.encode_with_indexes(k["symbols"][:3], k["indexes"][:3], t["cdfs"], t["sizes"], t["offsets"])
    e.encode_with_indexes(k["symbols"][3:], k["indexes"][3:], t["cdfs"], t["sizes"], t["offsets"])
    assert e.flush().hex() == k["hex"]
    # reusable after flush
    e.encode_with_indexes(k["symbols"], k["indexes"], t["cdfs"], t["sizes"], t["offsets"])
    assert e.flush().hex() == k["hex"]
    with pytest.raises(ValueError):
        ans.RansEncoder().encode_with_indexes([0], [7], t["cdfs"], t["sizes"], t["offsets"])


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


def test_seeded_streams_match_reference_binary(env):
    ans = env["ans"]
    for k in env["kat"]["streams"]:
        sym, idx = _stream(k, env)
        s = torch.from_numpy(sym).cuda().unsqueeze(0)
        i = torch.from_numpy(idx).cuda().unsqueeze(0)
        b = ans.encode_streams(env["tables"], s, i)[0]
        assert len(b) == k["nbytes"], k
        assert hashlib.sha1(b).hexdigest() == k["sha1"], k
        # decode in three steps on one set_stream, like the 12 decode_stream calls of stf.py:751-766
        d = ans.StreamDecoder(1)
        d.set_streams([b])
        n = k["n"]
        parts = [d.decode_step(env["tables"], i[:, a:e]) for a, e in ((0, n // 5), (n // 5, n // 2), (n // 2, n))]
        d.check_status()
        assert torch.equal(torch.cat(parts, 1), s)


def test_many_ragged_streams_in_one_launch(env):
    """64 independent streams with different content; each must equal the oracle's one-at-a-time result."""
    ans, coder = env["ans"], env["coder"]
    cdf, lengths, offsets = env["gc"]
    rng = np.random.default_rng(99)
    S, N = 64, 3001
    idx = rng.integers(0, 64, (S, N)).astype(np.int32)
    idx[::7] = np.minimum(rng.geometric(0.3, (len(idx[::7]), N)) - 1, 63)
    sym = np.rint(rng.normal(0, env["table"][idx] * rng.uniform(0.2, 3.0, (S, 1)))).astype(np.int32)
    sym[3] = 0
    out = ans.encode_streams(env["tables"], torch.from_numpy(sym).cuda(), torch.from_numpy(idx).cuda())
    for s in range(S):
        assert out[s] == coder.rans_encode(sym[s], idx[s], cdf, lengths, offsets), s
    d = ans.StreamDecoder(S)
    d.set_streams(out)
    got = d.decode_step(env["tables"], torch.from_numpy(idx).cuda())
    assert np.array_equal(got.cpu().numpy(), sym)


def test_edge_cases(env):
    ans, coder = env["ans"], env["coder"]
    cdf, lengths, offsets = env["gc"]
    # 1 and 2 symbols (the reference's flush under-allocates here), all tables, extreme symbols
    for sym, idx in [([0], [0]), ([3, -3], [10, 63]), ([2 ** 26, -(2 ** 26), 0], [0, 1, 2]), (list(range(-40, 41)), [5] * 81)]:
        s = torch.tensor([sym], dtype=torch.int32).cuda()
        i = torch.tensor([idx], dtype=torch.int32).cuda()
        b = ans.encode_streams(env["tables"], s, i)[0]
        assert b == coder.rans_encode(sym, idx, cdf, lengths, offsets)
        assert ans.RansDecoder().decode_with_indexes(b, idx, env["tables"]) == sym
    assert len(ans.BufferedRansEncoder().flush()) == 8
    with pytest.raises(ValueError):
        ans.RansDecoder().set_stream(b"123")  # not a multiple of 4 bytes: StreamDecoder refuses
    with pytest.raises(ValueError):
        ans.RansDecoder().decode_with_indexes(b, [64], env["tables"])


def test_full_size_round_trip_property(env):
    """BASELINE size (589 824 symbols/stream, 8 streams): size-independent checks -- decode(encode(x)) == x,
    stream lengths multiples of 4, first 8 bytes = a normalised final state -- plus one stream vs the oracle."""
    ans, coder = env["ans"], env["coder"]
    cdf, lengths, offsets = env["gc"]
    rng = np.random.default_rng(5)
    S, N = 8, 589824
    idx = np.minimum(rng.geometric(0.08, (S, N)) - 1, 63).astype(np.int32)
    sym = np.rint(rng.normal(0, env["table"][idx])).astype(np.int32)
    ds, di = torch.from_numpy(sym).cuda(), torch.from_numpy(idx).cuda()
    out = ans.encode_streams(env["tables"], ds, di)
    for b in out:
        assert len(b) % 4 == 0
        state = int.from_bytes(b[:8], "little")
        assert (1 << 31) <= state < (1 << 63)
    assert out[0] == coder.rans_encode(sym[0], idx[0], cdf, lengths, offsets)
    d = ans.StreamDecoder(S)
    d.set_streams(out)
    steps = [d.decode_step(env["tables"], di[:, k * 49152:(k + 1) * 49152]) for k in range(12)]
    assert torch.equal(torch.cat(steps, 1), ds)
