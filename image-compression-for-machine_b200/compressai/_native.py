"""ctypes binding of libicm_b200.so (the C-ABI CUDA library, include/icm_b200.h).

There is NO fallback: if the library is missing or a call fails, an exception is raised.  PyTorch is
used only for device memory (tensor.data_ptr()) and the current CUDA stream.
"""
import ctypes as C
import os

import torch

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_PKG, "lib", "libicm_b200.so")

ACT_NONE, ACT_GELU, ACT_HALF_TANH, ACT_SIGMOID, ACT_RSQRT, ACT_SQRT = 0, 1, 2, 3, 4, 5
RES_ADD, RES_ADD_BEFORE_ACT, RES_MUL = 0, 1, 2
OUT_BF16, OUT_F32 = 0, 1
EB_PARAMS = 59


class NativeError(RuntimeError):
    pass


class View(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("sb", C.c_int64), ("sc", C.c_int64), ("sp", C.c_int64)]


class ConvArgs(C.Structure):
    _fields_ = [
        ("inp", C.c_void_p), ("weight", C.c_void_p), ("bias", C.c_void_p), ("out", C.c_void_p), ("residual", C.c_void_p),
        ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("Cin", C.c_int), ("in_pitch", C.c_int),
        ("Cout", C.c_int), ("out_pitch", C.c_int), ("KH", C.c_int), ("KW", C.c_int), ("stride", C.c_int), ("pad", C.c_int),
        ("act", C.c_int), ("out_dtype", C.c_int), ("pixel_shuffle", C.c_int), ("res_pitch", C.c_int),
        ("res_dtype", C.c_int), ("res_mode", C.c_int),
    ]


class Rows(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("stride", C.c_int64)]


class ConvGroups(C.Structure):
    _fields_ = [("groups", C.c_int), ("in_images", C.c_int), ("weight_group_rows", C.c_int64), ("bias_group_stride", C.c_int64),
                ("out_group_stride", C.c_int64), ("in_image_offset", C.c_int * 16), ("tail_channel", C.c_int * 16)]


_lib = None
_profile = None  # a Profile instance while per-call CUDA-event timing is switched on (bench.py)


class Profile:
    """Wraps every device entry point with a CUDA-event pair on the current stream (bench.py's roofline
    pass).  Usage: `with Profile() as p: model.compress(x)`; `p.summary()` -> {entry: (calls, ms, work)}."""

    _HOST = {"icm_last_error", "icm_abi_version", "icm_launch_count", "icm_note_graph_launches", "icm_pmf_to_quantized_cdf", "icm_tables_create",
             "icm_tables_destroy", "icm_rans_encode_workspace_bytes", "icm_rans_decoder_create", "icm_rans_decoder_destroy",
             "icm_rans_decoder_set_streams", "icm_rans_decoder_status", "icm_set_conv_sm_limit", "icm_set_decoder_streams_per_cta", "icm_set_decoder_layout"}

    def __init__(self):
        self.records = {}
        self._wrapped = {}

    def __enter__(self):
        global _profile
        self._L = _load()
        _profile = self
        return self

    def __exit__(self, *exc):
        global _profile
        _profile = None

    def __getattr__(self, name):
        fn = getattr(self._L, name)
        if name in self._HOST or not name.startswith("icm_"):
            return fn
        w = self._wrapped.get(name)
        if w is None:
            def w(*args, _fn=fn, _name=name):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                rc = _fn(*args)
                e.record()
                self.records.setdefault(_name, []).append((s, e, self._work(_name, args)))
                return rc
            self._wrapped[name] = w
        return w

    @staticmethod
    def _work(name, args):
        if name in ("icm_conv2d", "icm_conv2d_grouped"):  # algorithmic FLOPs: 2 * output pixels * Cout * taps * Cin (x groups)
            a = args[0]._obj
            Ho = (a.H + 2 * a.pad - a.KH) // a.stride + 1
            Wo = (a.W + 2 * a.pad - a.KW) // a.stride + 1
            G = args[1]._obj.groups if name == "icm_conv2d_grouped" else 1
            return 2.0 * G * a.B * Ho * Wo * a.Cout * a.KH * a.KW * a.Cin
        if name == "icm_swin_mlp":  # two products rows x 4C x C
            rows, Cc = args[6], args[7]
            return 2.0 * 2.0 * rows * 4 * Cc * Cc
        if name == "icm_swin_block":  # qkv + proj (4 C^2) and the MLP (8 C^2) per token, plus the 16x16x16 attention products
            rows, Cc, parts = args[1] * args[2] * args[3], args[4], args[8]
            return 2.0 * rows * Cc * ((4 * Cc + 2 * 16) * (parts & 1) + 8 * Cc * ((parts >> 1) & 1))
        return 0.0

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, recs in self.records.items():
            out[name] = (len(recs), sum(s.elapsed_time(e) for s, e, _ in recs), sum(w for _, _, w in recs))
        return out


def lib():
    """The loaded library (or the profiling proxy while a Profile is active)."""
    return _profile if _profile is not None else _load()


def _load():
    """Load the library (building it first if nvcc is available and it is absent)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        try:
            import importlib.util

            spec = importlib.util.spec_from_file_location("_icm_build", os.path.join(_PKG, "build.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mod.build()
        except Exception as e:  # noqa: BLE001
            raise NativeError(
                f"{LIB_PATH} is missing and could not be built ({e}); run `python image-compression-for-machine_b200/build.py`. "
                "There is no CPU fallback for this path."
            ) from e
    L = C.CDLL(LIB_PATH)
    P, I, I64, F = C.c_void_p, C.c_int, C.c_int64, C.c_float
    sig = {
        "icm_last_error": (C.c_char_p, []),
        "icm_abi_version": (I, []),
        "icm_launch_count": (I64, []),
        "icm_note_graph_launches": (I64, [I64]),
        "icm_pmf_to_quantized_cdf": (I, [P, I, I, P]),
        "icm_tables_create": (I, [P, I, I, P, P, C.POINTER(P)]),
        "icm_tables_destroy": (None, [P]),
        "icm_rans_encode_workspace_bytes": (I64, [I, I64]),
        "icm_rans_encode_batch": (I, [P, P, P, I, I64, P, P, I64, P, P]),
        "icm_rans_decoder_create": (I, [I, C.POINTER(P)]),
        "icm_rans_decoder_destroy": (None, [P]),
        "icm_rans_decoder_set_streams": (I, [P, P, P, P, P]),
        "icm_rans_decoder_set_streams_device": (I, [P, P, P, P]),
        "icm_set_decoder_streams_per_cta": (I, [I]),
        "icm_set_decoder_layout": (I, [I, I]),
        "icm_rans_decoder_step": (I, [P, P, P, I64, P, P]),
        "icm_rans_decoder_status": (I, [P, P, P]),
        "icm_rans_decoder_status_min": (I, [P, P, P]),
        "icm_min_i32": (I, [P, I64, P, P]),
        "icm_gc_quantize_index": (I, [View, View, View, I, I, I64, P, I, F, P, P, I64, I64, View, View, View, P]),
        "icm_gc_build_indexes": (I, [View, I, I, I64, P, I, F, P, I64, I64, P]),
        "icm_gc_dequantize": (I, [P, I64, I64, View, I, I, I64, View, View, View, P]),
        "icm_gc_likelihood": (I, [View, View, View, I, I, I64, F, F, View, View, View, View, P]),
        "icm_add_lrp": (I, [View, View, I, I, I64, View, View, P]),
        "icm_eb_process": (I, [I, View, I, I, I64, P, F, P, P, View, View, View, P]),
        "icm_conv2d": (I, [C.POINTER(ConvArgs), P]),
        "icm_conv2d_grouped": (I, [C.POINTER(ConvArgs), C.POINTER(ConvGroups), P]),
        "icm_swin_mlp": (I, [P, P, P, P, P, P, I64, I, P]),
        "icm_swin_block": (I, [P, I, I, I, I, I, I, I, I] + [P] * 14),
        "icm_set_conv_sm_limit": (I, [I]),
        "icm_pack_conv_weight": (I, [P, I, I, I, I, I, I, I, P, P]),
        "icm_layernorm": (I, [P, P, P, P, I, I64, I, I, I, I, I, P]),
        "icm_cast_bf16": (I, [P, I64, I, I64, P, I64, P]),
        "icm_window_attention": (I, [P, P, P, I, I, I, I, I, I, I, P]),
        "icm_patch_embed": (I, [P, P, P, P, P, P, I, I, I, I, P]),
        "icm_final_conv": (I, [P, P, P, P, I, I, I, I, I, P]),
        "icm_image_to_nhwc": (I, [P, P, I, I, I, I, I, P]),
        "icm_nhwc_to_image": (I, [P, P, I, I, I, I, I, I, P]),
        "icm_eltwise_bf16": (I, [I, P, I64, P, I64, P, I64, P, I64, I64, I, P]),
        "icm_window_attention_wacnn": (I, [P, P, P, I, I, I, I, I, I, I, P]),
        "icm_pack_deconv_weight": (I, [P, I, I, I, I, P, P]),
        "icm_gc_train_forward": (I, [Rows, Rows, Rows, Rows, I64, I64, F, F, Rows, Rows, P]),
        "icm_gc_train_backward": (I, [Rows, Rows, Rows, Rows, Rows, Rows, I64, I64, F, F, Rows, Rows, Rows, P]),
        "icm_layernorm_train_forward": (I, [P, P, P, P, I, P, P, I64, I, P]),
        "icm_layernorm_train_backward": (I, [P, P, I, P, P, P, P, P, P, I64, I, P]),
        "icm_grad_sumsq": (I, [P, I64, P, P]),
        "icm_clip_coef": (I, [P, F, F, P, P, P]),
        "icm_adam_step": (I, [P, P, P, P, I64, F, F, F, F, I, P, P, F, P]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    L._icm_symbols = tuple(sig)
    _lib = L
    return L


def check(rc, what=""):
    if rc is not None and rc < 0:
        msg = _load().icm_last_error().decode(errors="replace")
        raise NativeError(f"{what} failed ({rc}): {msg}")
    return rc


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr():
    """The current CUDA stream as a void*.  Called once per kernel launch: the raw-stream query is ~10x cheaper than building a
    torch.cuda.Stream object (it was 10 % of the host time of a pipelined step)."""
    if _raw_stream is not None:
        return C.c_void_p(_raw_stream(torch.cuda.current_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(t, name="tensor"):
    if not t.is_cuda:
        raise NativeError(f"{name} must live on a CUDA device: this path has no CPU implementation")


NULL_VIEW = View(None, 0, 0, 0)


def view_bcp(t, B, Cc, P, channel_offset=0):
    """View of a contiguous channels-last [B, P, pitch] tensor restricted to channels [off, off+C)."""
    if t is None:
        return NULL_VIEW
    pitch = t.shape[-1]
    return View(t.data_ptr() + channel_offset * t.element_size(), P * pitch, 1, pitch)


def view_nchw(t):
    """View of a contiguous [B, C, *spatial] tensor."""
    if t is None:
        return NULL_VIEW
    B, Cc = t.shape[0], t.shape[1]
    P = t[0, 0].numel()
    return View(t.data_ptr(), Cc * P, P, 1)


def launch_count():
    return int(_load().icm_launch_count())
