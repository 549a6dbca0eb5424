"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, and its host-only entry points agree with the fixtures recorded from the reference binaries."""
import ctypes as C
import hashlib
import json
import os
import re

import numpy as np
import pytest
import torch

import compressai
from compressai import _CXX, _native

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(REPO, "include", "icm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(icm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = _native.lib()
    names = _header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/icm_b200.h but not exported by libicm_b200.so"
    assert L.icm_abi_version() == 1
    # the python binding table covers the header too
    assert set(names) == set(L._icm_symbols)


def test_pmf_to_quantized_cdf_matches_reference_binary(golden_dir):
    kat = json.load(open(os.path.join(golden_dir, "rans_kat.json")))
    for k in kat["pmf"]:
        assert _CXX.pmf_to_quantized_cdf(k["pmf"], k["precision"]) == k["cdf"]
    with pytest.raises(TypeError):
        _CXX.pmf_to_quantized_cdf([0.5, 0.5], 16.0)
    with pytest.raises(_native.NativeError):
        _CXX.pmf_to_quantized_cdf([0.0, 0.0], 16)


def test_table_construction_matches_reference(golden_dir):
    """GaussianConditional.update()/EntropyBottleneck.update() are host code: exact tables without a GPU."""
    from compressai.entropy_models import EntropyBottleneck, GaussianConditional
    from compressai.models.stf import get_scale_table
    from oracle import weights

    kat = json.load(open(os.path.join(golden_dir, "rans_kat.json")))
    g = np.load(os.path.join(golden_dir, "gc_tables.npz"))
    gc = GaussianConditional(None)
    assert gc.update_scale_table(get_scale_table())
    assert not gc.update_scale_table(get_scale_table())  # already initialised, force=False
    assert np.array_equal(gc.cdf_length.numpy(), g["cdf_length"]) and np.array_equal(gc.offset.numpy(), g["offset"])
    assert hashlib.sha1(gc.quantized_cdf.numpy().astype("<i4").tobytes()).hexdigest() == kat["gc_table_sha1"]
    assert abs(-gc._standardized_quantile(0.5e-9) - 6.109410204869) < 1e-9

    e = np.load(os.path.join(golden_dir, "entropy_kat.npz"))
    eb = EntropyBottleneck(192)
    sd = weights.seeded_state_dict({k: v for k, v in eb.state_dict().items()}, seed=3, stress=False)
    eb.load_state_dict({k: v for k, v in sd.items() if k in dict(eb.named_parameters())}, strict=False)
    assert eb.update(force=True) and not eb.update()
    assert np.array_equal(eb.quantized_cdf.numpy(), e["eb_cdf"])
    assert np.array_equal(eb.cdf_length.numpy(), e["eb_len"]) and np.array_equal(eb.offset.numpy(), e["eb_off"])


def test_registry_and_errors():
    assert compressai.get_entropy_coder() == "ans" and compressai.available_entropy_coders() == ["ans"]
    with pytest.raises(ValueError):
        compressai.set_entropy_coder("rangecoder")
    from compressai.entropy_models import EntropyBottleneck, GaussianConditional

    with pytest.raises(ValueError):
        GaussianConditional([3.0, 1.0])
    with pytest.raises(ValueError):
        GaussianConditional(None, scale_bound=-1)
    eb = EntropyBottleneck(8)
    with pytest.raises(ValueError):
        eb.quantize(torch.zeros(1, 8, 2, 2), "bogus")
    with pytest.raises(ValueError, match="Run update"):
        eb.device_tables()
    # no silent CPU fallback: CPU tensors are refused loudly
    with pytest.raises(_native.NativeError):
        eb.quantize(torch.zeros(1, 8, 2, 2), "symbols")
    assert eb._build_indexes((2, 8, 3, 3)).shape == (2, 8, 3, 3) and int(eb._build_indexes((2, 8, 3, 3))[1, 5, 2, 2]) == 5


def test_model_state_dict_names_match_reference_layout():
    from compressai.zoo import models
    from oracle import stf_ref

    m = models["stf"]()
    mine = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    ref = {k: tuple(v.shape) for k, v in stf_ref.template_state_dict().items()}
    assert len(mine) == 779
    for k, s in ref.items():
        assert mine.get(k) == s, k
    with pytest.raises(_native.NativeError):
        m.eval()(torch.zeros(1, 3, 64, 64))  # CPU tensor: refused, not computed on the host
