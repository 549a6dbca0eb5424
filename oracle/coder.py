"""ctypes front-end of oracle/rans_oracle.c (TEST INFRASTRUCTURE -- see that file's header).

Mirrors the reference's `compressai.ans` / `compressai._CXX` call shapes on numpy arrays:
  pmf_to_quantized_cdf  <- /root/reference/compressai/entropy_models/entropy_models.py:60-63
  RansEncoder / BufferedRansEncoder / RansDecoder
                        <- /root/reference/compressai/models/stf.py:698,727-729,751-752,766
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")


def build(force=False):
    src = os.path.join(_HERE, "rans_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        i32p, u8p, i64p = C.POINTER(C.c_int32), C.POINTER(C.c_uint8), C.POINTER(C.c_int64)
        L.orc_pmf_to_quantized_cdf.argtypes = [C.POINTER(C.c_float), C.c_int, C.c_int, C.POINTER(C.c_uint32)]
        L.orc_pmf_to_quantized_cdf.restype = C.c_int
        tbl = [i32p, C.c_int, C.c_int, i32p, i32p]
        L.orc_rans_encode.argtypes = [i32p, i32p, C.c_int64] + tbl + [u8p, C.c_int64]
        L.orc_rans_encode.restype = C.c_int64
        L.orc_rans_decode.argtypes = [u8p, C.c_int64, i32p, C.c_int64] + tbl + [i32p, i64p]
        L.orc_rans_decode.restype = C.c_int
        L.orc_encoder_new.restype = C.c_void_p
        L.orc_encoder_free.argtypes = [C.c_void_p]
        L.orc_encoder_push.argtypes = [C.c_void_p, i32p, i32p, C.c_int64] + tbl
        L.orc_encoder_push.restype = C.c_int
        L.orc_encoder_flush.argtypes = [C.c_void_p, u8p, C.c_int64]
        L.orc_encoder_flush.restype = C.c_int64
        L.orc_encoder_pending.argtypes = [C.c_void_p]
        L.orc_encoder_pending.restype = C.c_int64
        _lib = L
    return _lib


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _tables(cdfs, sizes, offsets):
    cdfs = _i32(cdfs)
    assert cdfs.ndim == 2
    sizes, offsets = _i32(sizes), _i32(offsets)
    return cdfs, sizes, offsets, [_p(cdfs, C.c_int32), cdfs.shape[0], cdfs.shape[1], _p(sizes, C.c_int32), _p(offsets, C.c_int32)]


def pmf_to_quantized_cdf(pmf, precision=16):
    pmf = np.ascontiguousarray(pmf, dtype=np.float32)
    out = np.zeros(pmf.size + 1, np.uint32)
    rc = lib().orc_pmf_to_quantized_cdf(_p(pmf, C.c_float), pmf.size, precision, _p(out, C.c_uint32))
    if rc < 0:
        raise ValueError(f"pmf_to_quantized_cdf failed ({rc})")
    return out.astype(np.int64)


def _cap(n):
    return 8 * int(n) * 4 + 64  # generous: <= 1 word per coder record, few records per symbol


def rans_encode(symbols, indexes, cdfs, sizes, offsets):
    symbols, indexes = _i32(symbols).ravel(), _i32(indexes).ravel()
    assert symbols.size == indexes.size
    cdfs, sizes, offsets, t = _tables(cdfs, sizes, offsets)
    cap = _cap(symbols.size)
    while True:
        out = np.empty(cap, np.uint8)
        nb = lib().orc_rans_encode(_p(symbols, C.c_int32), _p(indexes, C.c_int32), symbols.size, *t, _p(out, C.c_uint8), cap)
        if nb == -4:
            cap *= 4
            continue
        if nb < 0:
            raise ValueError(f"rans_encode failed ({nb})")
        return out[:nb].tobytes()


class RansEncoder:
    def encode_with_indexes(self, symbols, indexes, cdfs, sizes, offsets):
        return rans_encode(symbols, indexes, cdfs, sizes, offsets)


class BufferedRansEncoder:
    def __init__(self):
        self._h = lib().orc_encoder_new()

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_encoder_free(self._h)
            self._h = None

    def encode_with_indexes(self, symbols, indexes, cdfs, sizes, offsets):
        symbols, indexes = _i32(symbols).ravel(), _i32(indexes).ravel()
        cdfs, sizes, offsets, t = _tables(cdfs, sizes, offsets)
        rc = lib().orc_encoder_push(self._h, _p(symbols, C.c_int32), _p(indexes, C.c_int32), symbols.size, *t)
        if rc:
            raise ValueError(f"encode_with_indexes failed ({rc})")

    def flush(self):
        cap = int(lib().orc_encoder_pending(self._h)) * 4 + 16
        out = np.empty(cap, np.uint8)
        nb = lib().orc_encoder_flush(self._h, _p(out, C.c_uint8), cap)
        if nb < 0:
            raise ValueError(f"flush failed ({nb})")
        return out[:nb].tobytes()


class RansDecoder:
    def __init__(self):
        self._stream = None
        self._state = np.array([0, -1], np.int64)

    def set_stream(self, data):
        self._stream = np.frombuffer(bytes(data), np.uint8).copy()
        self._state[:] = (0, -1)

    def decode_stream(self, indexes, cdfs, sizes, offsets):
        indexes = _i32(indexes).ravel()
        cdfs, sizes, offsets, t = _tables(cdfs, sizes, offsets)
        out = np.empty(indexes.size, np.int32)
        rc = lib().orc_rans_decode(
            _p(self._stream, C.c_uint8), self._stream.size, _p(indexes, C.c_int32), indexes.size, *t,
            _p(out, C.c_int32), _p(self._state, C.c_int64),
        )
        if rc:
            raise ValueError(f"decode_stream failed ({rc})")
        return out

    def decode_with_indexes(self, data, indexes, cdfs, sizes, offsets):
        self.set_stream(data)
        return self.decode_stream(indexes, cdfs, sizes, offsets)
