"""`compressai.ans` replacement: the reference's rANS coder running on the GPU.

Reference interface (pybind11 module compressai/ans, binary only -- ans.cpython-38-x86_64-linux-gnu.so;
signatures from its embedded docstrings; call sites entropy_models.py:228,277 and
models/stf.py:698,727-729,751-752,766):

    RansEncoder().encode_with_indexes(symbols, indexes, cdfs, cdfs_sizes, offsets) -> bytes
    BufferedRansEncoder().encode_with_indexes(symbols, indexes, cdfs, cdfs_sizes, offsets) -> None
    BufferedRansEncoder().flush() -> bytes
    RansDecoder().set_stream(bytes) -> None
    RansDecoder().decode_stream(indexes, cdfs, cdfs_sizes, offsets) -> list[int]
    RansDecoder().decode_with_indexes(bytes, indexes, cdfs, cdfs_sizes, offsets) -> list[int]

The classes below keep those names, argument meanings and results (byte-identical streams), and accept
Python lists exactly like the reference.  They also accept CUDA int32 tensors and a `Tables` object in
place of `cdfs` (then `cdfs_sizes` / `offsets` may be None), which is what the in-package entropy models
use so that nothing round-trips through Python lists.  `encode_streams` / `StreamDecoder` are the batched
forms (one independent rANS state per stream, many streams per launch).

Differences from the reference, all on inputs where the reference has undefined behaviour:
an out-of-range CDF index raises ValueError (reference: compiled-out assert), flush() of an empty encoder
returns the 8-byte initial state (reference: NULL dereference), |symbol - offset| must be < 2**27.
"""
import ctypes as C

import numpy as np
import torch

from . import _native
from ._native import NativeError, check, lib, stream_ptr

_ERR_BAD_INDEX = -5
_ERR_CAPACITY = -4


def _device():
    if not torch.cuda.is_available():
        raise NativeError("compressai.ans needs a CUDA device: the rANS coder has no CPU implementation in this package")
    return torch.device("cuda", torch.cuda.current_device())


def _as_i32_cuda(x, name):
    if isinstance(x, torch.Tensor):
        t = x
        if t.dtype != torch.int32:
            t = t.to(torch.int32)
        if not t.is_cuda:
            t = t.to(_device(), non_blocking=True)
        return t.contiguous().reshape(-1)
    try:
        arr = np.asarray(x, dtype=np.int32).reshape(-1)
    except (TypeError, ValueError) as e:
        raise TypeError(f"{name}: expected a list of ints") from e
    return torch.from_numpy(arr).to(_device())


class Tables:
    """Quantised CDF tables resident on the device (icm_tables).  Build once, reuse for every call."""

    def __init__(self, cdfs, cdfs_sizes, offsets):
        to_np = lambda v: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)).astype(np.int32)
        if not isinstance(cdfs, (torch.Tensor, np.ndarray)):
            width = max(len(r) for r in cdfs)
            rows = np.zeros((len(cdfs), width), np.int32)
            for i, r in enumerate(cdfs):
                rows[i, : len(r)] = r
            cdfs = rows
        self.cdfs = np.ascontiguousarray(to_np(cdfs))
        self.sizes = np.ascontiguousarray(to_np(cdfs_sizes).reshape(-1))
        self.offsets = np.ascontiguousarray(to_np(offsets).reshape(-1))
        if self.cdfs.ndim != 2 or self.sizes.size != self.cdfs.shape[0] or self.offsets.size != self.cdfs.shape[0]:
            raise ValueError("cdfs must be [n, L] with n sizes and n offsets")
        _device()
        h = C.c_void_p()
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        check(lib().icm_tables_create(p(self.cdfs), self.cdfs.shape[0], self.cdfs.shape[1], p(self.sizes), p(self.offsets), C.byref(h)),
              "icm_tables_create")
        self.handle = h
        self.n_cdf = self.cdfs.shape[0]

    def __del__(self):
        h = getattr(self, "handle", None)
        L = getattr(_native, "_lib", None) if _native is not None else None  # module globals vanish at interpreter exit
        if h and L is not None:
            L.icm_tables_destroy(h)
            self.handle = None


_table_cache = {}


def _tables_from_args(cdfs, cdfs_sizes, offsets):
    """Device tables for the reference-style (cdfs, cdfs_sizes, offsets) arguments.  The reference converts the whole
    list-of-lists on every call; here the lists are flattened once per call and the device copy is reused whenever
    the CONTENT is unchanged (keyed on a digest, so lists mutated in place are never served stale tables)."""
    if isinstance(cdfs, Tables):
        return cdfs
    import hashlib

    to_np = lambda v: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)).astype(np.int32)
    if isinstance(cdfs, (torch.Tensor, np.ndarray)):
        rows = np.ascontiguousarray(to_np(cdfs))
    else:
        width = max(len(r) for r in cdfs)
        rows = np.zeros((len(cdfs), width), np.int32)
        for i, r in enumerate(cdfs):
            rows[i, : len(r)] = r
    sizes, offs = np.ascontiguousarray(to_np(cdfs_sizes).reshape(-1)), np.ascontiguousarray(to_np(offsets).reshape(-1))
    h = hashlib.blake2b(digest_size=16)
    for a in (np.asarray(rows.shape, np.int64), rows, sizes, offs):
        h.update(a.tobytes())
    key = (torch.cuda.current_device() if torch.cuda.is_available() else -1, h.digest())
    hit = _table_cache.get(key)
    if hit is None:
        hit = Tables(rows, sizes, offs)
        if len(_table_cache) > 16:
            _table_cache.clear()
        _table_cache[key] = hit
    return hit


_workspaces = {}


def _workspace(nbytes, device):
    key = (device.index, torch.cuda.current_stream().cuda_stream)  # one per stream: micro-batches run concurrently
    w = _workspaces.get(key)
    if w is None or w.numel() < nbytes:
        w = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=device)
        _workspaces[key] = w
    return w


def raise_for_status(code, what="rANS coder"):
    """Map a (folded) device status word to the exception the host-string path raises."""
    if code == _ERR_BAD_INDEX:
        raise ValueError(f"{what}: CDF index out of range")
    if code == _ERR_CAPACITY:
        raise CapacityError(f"{what}: bit-stream larger than the optimistic 16 bit/symbol buffer")
    if code < 0:
        raise NativeError(f"{what}: device status {code}")


class CapacityError(NativeError):
    """The asynchronous encoder's optimistic output buffer was too small; retry with worst_case=True."""


def fold_status(values, flag):
    """flag (CUDA int32[1]) = min(flag, values): asynchronous, no host synchronisation."""
    check(lib().icm_min_i32(values.data_ptr(), values.numel(), flag.data_ptr(), stream_ptr()), "icm_min_i32")


def encode_streams(tables, symbols, indexes, return_device=False, worst_case=False):
    """Encode S independent streams.  symbols / indexes: CUDA int32 [S, N] in stream order.

    Returns a list of S `bytes` (or, with return_device=True, (packed uint8 CUDA tensor, sizes list)).
    return_device="async": (packed, sizes int32[S+1]) with no host synchronisation; a negative size is an ICM_ERR_* code
    (the output buffer holds 16 bit/symbol unless worst_case=True: 7 bytes/symbol)."""
    assert symbols.is_cuda and indexes.is_cuda and symbols.dtype == torch.int32 and indexes.dtype == torch.int32
    S, N = symbols.shape
    assert indexes.shape == symbols.shape
    symbols, indexes = symbols.contiguous(), indexes.contiguous()
    dev = symbols.device
    L = lib()
    work = _workspace(L.icm_rans_encode_workspace_bytes(S, N), dev)
    sizes = torch.empty(S + 1, dtype=torch.int32, device=dev)
    cap = S * (7 * N + 256) if worst_case else S * (2 * N + 64)  # 16 bit/symbol: ample for model data; retried at the worst case if exceeded
    if return_device == "async":  # no host synchronisation: caller inspects `sizes` (negative = error code) later
        packed = torch.empty(cap, dtype=torch.uint8, device=dev)
        check(L.icm_rans_encode_batch(tables.handle, symbols.data_ptr(), indexes.data_ptr(), S, N, work.data_ptr(),
                                      packed.data_ptr(), cap, sizes.data_ptr(), stream_ptr()), "icm_rans_encode_batch")
        return packed, sizes
    for attempt in range(2):
        packed = torch.empty(cap, dtype=torch.uint8, device=dev)
        check(L.icm_rans_encode_batch(tables.handle, symbols.data_ptr(), indexes.data_ptr(), S, N, work.data_ptr(),
                                      packed.data_ptr(), cap, sizes.data_ptr(), stream_ptr()), "icm_rans_encode_batch")
        hs = sizes.cpu().tolist()
        if any(v == _ERR_BAD_INDEX for v in hs[:S]):
            raise ValueError("encode_with_indexes: CDF index out of range")
        if any(v == _ERR_CAPACITY for v in hs[:S]):
            if attempt == 0:
                cap = S * (7 * N + 256)
                continue
            raise NativeError("rANS output exceeded the worst-case capacity")
        break
    total = sum(hs[:S])
    if return_device:
        return packed[:total], hs[:S]
    host = packed[:total].cpu().numpy().tobytes()
    out, o = [], 0
    for v in hs[:S]:
        out.append(host[o:o + v])
        o += v
    return out


def strings_to_host(packed, sizes, retry=None):
    """Device-resident streams (packed uint8, int32 sizes[S+1]) -> list of S `bytes`.  `retry()` re-encodes with
    the worst-case capacity when the optimistic output buffer of the asynchronous encode was too small."""
    hs = sizes.cpu().tolist()
    S = len(hs) - 1
    if any(v == _ERR_BAD_INDEX for v in hs[:S]):
        raise ValueError("encode_with_indexes: CDF index out of range")
    if any(v < 0 for v in hs[:S]):
        if retry is None:
            raise NativeError("rANS output exceeded the buffer capacity")
        return retry()
    total = sum(hs[:S])
    host = packed[:total].cpu().numpy().tobytes()
    out, o = [], 0
    for v in hs[:S]:
        out.append(host[o:o + v])
        o += v
    return out


_decoder_pool = {}


def acquire_decoder(n_streams):
    """A StreamDecoder from a per-size pool.  Creating / destroying one costs a cudaMalloc / cudaFree, and
    cudaFree synchronises the whole device, which would serialise concurrently running micro-batches."""
    free = _decoder_pool.setdefault((int(n_streams), torch.cuda.current_device()), [])
    return free.pop() if free else StreamDecoder(n_streams)


def release_decoder(dec):
    _decoder_pool.setdefault((dec.n_streams, torch.cuda.current_device()), []).append(dec)


class StreamDecoder:
    """S rANS decoders advancing in lock-step (icm_rans_decoder): set_streams once, then decode_step."""

    def __init__(self, n_streams):
        _device()
        h = C.c_void_p()
        check(lib().icm_rans_decoder_create(int(n_streams), C.byref(h)), "icm_rans_decoder_create")
        self.handle = h
        self.n_streams = int(n_streams)
        self._bytes = None

    def __del__(self):
        h = getattr(self, "handle", None)
        L = getattr(_native, "_lib", None) if _native is not None else None
        if h and L is not None:
            L.icm_rans_decoder_destroy(h)
            self.handle = None

    def set_streams(self, strings):
        if len(strings) != self.n_streams:
            raise ValueError("number of strings does not match the decoder")
        sizes = np.array([len(s) for s in strings], np.int64)
        if np.any(sizes % 4):
            raise ValueError("rANS streams are sequences of 32-bit words; got a length that is not a multiple of 4")
        offsets = np.zeros(self.n_streams, np.int64)
        offsets[1:] = np.cumsum(sizes)[:-1]
        blob = np.frombuffer(b"".join(bytes(s) for s in strings) + b"\0" * 4, np.uint8)
        self._bytes = torch.from_numpy(blob.copy()).to(_device())
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        check(lib().icm_rans_decoder_set_streams(self.handle, self._bytes.data_ptr(), p(offsets), p(sizes), stream_ptr()),
              "icm_rans_decoder_set_streams")

    def set_streams_device(self, packed, sizes):
        """Streams as `encode_streams(..., return_device="async")` left them on the device (no host copy)."""
        assert packed.is_cuda and sizes.is_cuda and sizes.dtype == torch.int32 and sizes.numel() >= self.n_streams
        self._bytes = packed
        self._sizes = sizes
        check(lib().icm_rans_decoder_set_streams_device(self.handle, packed.data_ptr(), sizes.data_ptr(), stream_ptr()),
              "icm_rans_decoder_set_streams_device")

    def decode_step(self, tables, indexes, out=None):
        """indexes: CUDA int32 [S, N] -> CUDA int32 [S, N] symbols."""
        assert indexes.is_cuda and indexes.dtype == torch.int32 and indexes.shape[0] == self.n_streams
        indexes = indexes.contiguous()
        if out is None:
            out = torch.empty_like(indexes)
        check(lib().icm_rans_decoder_step(self.handle, tables.handle, indexes.data_ptr(), indexes.shape[1], out.data_ptr(), stream_ptr()),
              "icm_rans_decoder_step")
        return out

    def fold_status(self, flag):
        """flag (CUDA int32[1]) = min(flag, this decoder's per-stream statuses); asynchronous."""
        check(lib().icm_rans_decoder_status_min(self.handle, flag.data_ptr(), stream_ptr()), "icm_rans_decoder_status_min")

    def check_status(self):
        st = np.zeros(self.n_streams, np.int32)
        check(lib().icm_rans_decoder_status(self.handle, st.ctypes.data_as(C.c_void_p), stream_ptr()), "icm_rans_decoder_status")
        if np.any(st == _ERR_BAD_INDEX):
            raise ValueError("decode: CDF index out of range")
        if np.any(st < 0):
            raise_for_status(int(st.min()), "decode (status handed over by the encoder)")


# ------------------------------------------------------------------------------------------------
# reference-shaped classes
class BufferedRansEncoder:
    def __init__(self):
        self._sym, self._idx, self._tables = [], [], None

    def encode_with_indexes(self, symbols, indexes, cdfs, cdfs_sizes=None, offsets=None):
        s, i = _as_i32_cuda(symbols, "symbols"), _as_i32_cuda(indexes, "indexes")
        if s.numel() != i.numel():
            raise ValueError("symbols and indexes must have the same length")
        t = _tables_from_args(cdfs, cdfs_sizes, offsets)
        if self._tables is not None and t is not self._tables and self._sym:
            raise NativeError("BufferedRansEncoder: all encode_with_indexes calls before a flush must use the same tables")
        self._tables = t
        self._sym.append(s)
        self._idx.append(i)

    def flush(self):
        if not self._sym:
            self._tables = None
            return (1 << 31).to_bytes(8, "little")  # the untouched initial state; the reference crashes here
        s, i = torch.cat(self._sym).unsqueeze(0), torch.cat(self._idx).unsqueeze(0)
        t = self._tables
        self._sym, self._idx, self._tables = [], [], None
        return encode_streams(t, s, i)[0]


class RansEncoder:
    def encode_with_indexes(self, symbols, indexes, cdfs, cdfs_sizes=None, offsets=None):
        e = BufferedRansEncoder()
        e.encode_with_indexes(symbols, indexes, cdfs, cdfs_sizes, offsets)
        return e.flush()


class RansDecoder:
    def __init__(self):
        self._dec = None

    def set_stream(self, encoded):
        if not isinstance(encoded, (bytes, bytearray, memoryview)):
            raise TypeError("set_stream(): expected bytes")
        self._dec = StreamDecoder(1)
        self._dec.set_streams([bytes(encoded)])

    def decode_stream(self, indexes, cdfs, cdfs_sizes=None, offsets=None, as_tensor=False):
        if self._dec is None:
            raise NativeError("decode_stream() before set_stream()")
        i = _as_i32_cuda(indexes, "indexes")
        t = _tables_from_args(cdfs, cdfs_sizes, offsets)
        out = self._dec.decode_step(t, i.unsqueeze(0))[0]
        self._dec.check_status()
        if as_tensor or isinstance(indexes, torch.Tensor):
            return out
        return out.cpu().tolist()

    def decode_with_indexes(self, encoded, indexes, cdfs, cdfs_sizes=None, offsets=None, as_tensor=False):
        self.set_stream(encoded)
        return self.decode_stream(indexes, cdfs, cdfs_sizes, offsets, as_tensor=as_tensor)
