"""`compressai._CXX` replacement: pmf_to_quantized_cdf through the C ABI (host code, no GPU needed).

Reference: pybind11 module compressai/_CXX (binary only, _CXX.cpython-38-x86_64-linux-gnu.so @0x68c0),
signature `pmf_to_quantized_cdf(list[float] pmf, int precision) -> list[int]`, called from
/root/reference/compressai/entropy_models/entropy_models.py:60-63.
"""
import ctypes as C

import numpy as np

from . import _native


def pmf_to_quantized_cdf(pmf, precision):
    if not isinstance(precision, int):
        raise TypeError("pmf_to_quantized_cdf(): incompatible function arguments (precision must be int)")
    arr = np.ascontiguousarray(np.asarray(pmf, dtype=np.float32).reshape(-1))
    out = np.empty(arr.size + 1, np.uint32)
    rc = _native.lib().icm_pmf_to_quantized_cdf(arr.ctypes.data_as(C.c_void_p), arr.size, precision, out.ctypes.data_as(C.c_void_p))
    _native.check(rc, "pmf_to_quantized_cdf")
    return out.tolist()
