"""Binary oracle: run the reference's OWN compiled rANS / CDF code under this interpreter.

TEST INFRASTRUCTURE ONLY (used by oracle/make_golden.py in the build container; it reads
/root/reference, which does not exist on the GPU box, so nothing in tests/, bench.py or
smoke() imports this module at run time).

The reference ships only cp38 binaries of its two native modules
(/root/reference/compressai/ans.cpython-38-x86_64-linux-gnu.so and _CXX.cpython-38-...so;
their sources, compressai/cpp_exts/*, are absent: /root/reference/setup.py:49-79).  The C++
functions inside are local, unstripped symbols; we load the .so with ctypes.PyDLL (lazy binding, so
the cp38-only CPython symbols are never resolved) and call the functions at their `nm` offsets with
hand-built libstdc++ std::vector / std::string objects.

This gives bit-exact outputs of the *reference implementation itself* for
  R1/R2  BufferedRansEncoder::encode_with_indexes / ::flush
  R3     RansEncoder::encode_with_indexes
  R4     RansDecoder::set_stream / ::decode_stream / ::decode_with_indexes
  R5     pmf_to_quantized_cdf
against which oracle/rans_oracle.c (the portable restatement) is pinned.
"""
import ctypes as C
import os

import numpy as np

REF = "/root/reference/compressai"
_ANS_SO = os.path.join(REF, "ans.cpython-38-x86_64-linux-gnu.so")
_CXX_SO = os.path.join(REF, "_CXX.cpython-38-x86_64-linux-gnu.so")

# offsets from `nm --defined-only -C`
_OFF = dict(
    PyInit_ans=0xA1A0,
    buf_encode=0x8A10,
    buf_flush=0x8730,
    enc_encode=0x8D70,
    dec_set_stream=0x7C40,
    dec_decode_stream=0x7CE0,
    dec_decode_with_indexes=0x8060,
    PyInit__CXX=0x6E90,
    pmf_to_cdf=0x68C0,
)


def available():
    return os.path.exists(_ANS_SO) and os.path.exists(_CXX_SO)


class _Vec(C.Structure):
    """libstdc++ std::vector<T>: {T* begin; T* end; T* end_of_storage}."""

    _fields_ = [("begin", C.c_void_p), ("end", C.c_void_p), ("cap", C.c_void_p)]


class _Str(C.Structure):
    """libstdc++ std::__cxx11::string: {char* p; size_t len; union {char buf[16]; size_t cap;}}."""

    _fields_ = [("p", C.c_void_p), ("len", C.c_size_t), ("buf", C.c_char * 16)]


class _Decoder(C.Structure):
    """RansDecoder: {uint64 state; std::string stream; uint32* ptr}."""

    _fields_ = [("state", C.c_uint64), ("stream", _Str), ("ptr", C.c_void_p)]


_keep = []  # keeps numpy buffers alive while the C++ side references them


def _vec_of(arr):
    arr = np.ascontiguousarray(arr)
    _keep.append(arr)
    v = _Vec()
    v.begin = arr.ctypes.data
    v.end = arr.ctypes.data + arr.nbytes
    v.cap = v.end
    return v


def _vecvec_of(rows):
    """std::vector<std::vector<int>> from a 2-D int32 array (every inner vector = one full row)."""
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    inner = (_Vec * rows.shape[0])()
    for i in range(rows.shape[0]):
        r = rows[i]
        inner[i].begin = r.ctypes.data
        inner[i].end = r.ctypes.data + r.nbytes
        inner[i].cap = inner[i].end
    _keep.append(rows)
    _keep.append(inner)
    v = _Vec()
    v.begin = C.addressof(inner)
    v.end = v.begin + C.sizeof(inner)
    v.cap = v.end
    return v


def _str_of(data: bytes):
    buf = C.create_string_buffer(data, len(data) + 1)
    _keep.append(buf)
    s = _Str()
    s.p = C.addressof(buf)
    s.len = len(data)
    return s


def _vec_to_np(v, dtype):
    n = (v.end - v.begin) // np.dtype(dtype).itemsize
    if n == 0:
        return np.zeros(0, dtype)
    return np.ctypeslib.as_array(C.cast(v.begin, C.POINTER(C.c_uint8)), (v.end - v.begin,)).view(dtype).copy()


class _Lib:
    def __init__(self):
        self.ans = C.PyDLL(_ANS_SO, mode=os.RTLD_LAZY)
        self.cxx = C.PyDLL(_CXX_SO, mode=os.RTLD_LAZY)
        self.ans_base = C.cast(self.ans.PyInit_ans, C.c_void_p).value - _OFF["PyInit_ans"]
        self.cxx_base = C.cast(self.cxx.PyInit__CXX, C.c_void_p).value - _OFF["PyInit__CXX"]
        P = C.c_void_p
        f = C.PYFUNCTYPE
        b = self.ans_base
        self.buf_encode = f(None, P, P, P, P, P, P)(b + _OFF["buf_encode"])
        self.buf_flush = f(P, P, P)(b + _OFF["buf_flush"])  # (sret, this) -> sret
        self.enc_encode = f(P, P, P, P, P, P, P, P)(b + _OFF["enc_encode"])  # (sret, this, 5 vecs)
        self.dec_set_stream = f(None, P, P)(b + _OFF["dec_set_stream"])
        self.dec_decode_stream = f(P, P, P, P, P, P, P)(b + _OFF["dec_decode_stream"])
        self.dec_decode_with_indexes = f(P, P, P, P, P, P, P, P)(b + _OFF["dec_decode_with_indexes"])
        self.pmf_to_cdf = f(P, P, P, C.c_int)(self.cxx_base + _OFF["pmf_to_cdf"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib


def _take_bytes(slot):
    """slot holds a new reference to a PyBytes object (py::bytes returned through sret)."""
    obj = C.cast(slot.value, C.py_object).value
    out = bytes(obj)
    C.pythonapi.Py_DecRef(C.py_object(obj))
    return out


def pmf_to_quantized_cdf(pmf, precision=16):
    L = lib()
    out = _Vec()
    v = _vec_of(np.asarray(pmf, dtype=np.float32))
    L.pmf_to_cdf(C.addressof(out), C.addressof(v), int(precision))
    return _vec_to_np(out, np.uint32).astype(np.int64)


def _args(symbols, indexes, cdfs, cdf_sizes, offsets):
    vs = []
    if symbols is not None:
        vs.append(_vec_of(np.asarray(symbols, dtype=np.int32)))
    vs.append(_vec_of(np.asarray(indexes, dtype=np.int32)))
    vs.append(_vecvec_of(cdfs))
    vs.append(_vec_of(np.asarray(cdf_sizes, dtype=np.int32)))
    vs.append(_vec_of(np.asarray(offsets, dtype=np.int32)))
    _keep.append(vs)
    return [C.addressof(v) for v in vs]


class BufferedRansEncoder:
    def __init__(self):
        self._obj = _Vec()  # an empty std::vector<RansSymbol>

    def encode_with_indexes(self, symbols, indexes, cdfs, cdf_sizes, offsets):
        lib().buf_encode(C.addressof(self._obj), *_args(symbols, indexes, cdfs, cdf_sizes, offsets))
        _keep.clear()

    def flush(self):
        slot = C.c_void_p()
        lib().buf_flush(C.addressof(slot), C.addressof(self._obj))
        return _take_bytes(slot)


class RansEncoder:
    def encode_with_indexes(self, symbols, indexes, cdfs, cdf_sizes, offsets):
        slot = C.c_void_p()
        this = C.c_uint64(0)
        lib().enc_encode(C.addressof(slot), C.addressof(this), *_args(symbols, indexes, cdfs, cdf_sizes, offsets))
        _keep.clear()
        return _take_bytes(slot)


class RansDecoder:
    def __init__(self):
        self._obj = _Decoder()
        self._obj.stream.p = C.addressof(self._obj.stream) + _Str.buf.offset  # empty SSO string
        self._obj.stream.len = 0

    def set_stream(self, data: bytes):
        s = _str_of(data)
        lib().dec_set_stream(C.addressof(self._obj), C.addressof(s))
        _keep.clear()

    def decode_stream(self, indexes, cdfs, cdf_sizes, offsets):
        out = _Vec()
        lib().dec_decode_stream(C.addressof(out), C.addressof(self._obj), *_args(None, indexes, cdfs, cdf_sizes, offsets))
        _keep.clear()
        return _vec_to_np(out, np.int32)

    def decode_with_indexes(self, data, indexes, cdfs, cdf_sizes, offsets):
        out = _Vec()
        s = _str_of(data)
        lib().dec_decode_with_indexes(
            C.addressof(out), C.addressof(self._obj), C.addressof(s), *_args(None, indexes, cdfs, cdf_sizes, offsets)
        )
        _keep.clear()
        return _vec_to_np(out, np.int32)


if __name__ == "__main__":
    # sanity KATs from SURVEY.md §8c
    print(pmf_to_quantized_cdf([0.25, 0.5, 0.25]))
    cdfs = np.array([[0, 16384, 49152, 65536, 0, 0], [0, 1, 32768, 65535, 65536, 0]], np.int32)
    s = RansEncoder().encode_with_indexes([-1, 0, -1, 0], [0, 0, 0, 0], cdfs, [4, 5], [-1, -1])
    print(s.hex(), "expect 0000090020000000")
    print(RansDecoder().decode_with_indexes(s, [0, 0, 0, 0], cdfs, [4, 5], [-1, -1]))
    e = BufferedRansEncoder()
    e.encode_with_indexes([100000, 0, 0, 0], [0, 0, 0, 0], cdfs, [4, 5], [-1, -1])
    print(e.flush().hex(), "expect e5d3c30000050010")
