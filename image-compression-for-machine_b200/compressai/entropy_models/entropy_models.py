"""Entropy models with the reference's API, running on the C-ABI CUDA kernels.

Mirrors /root/reference/compressai/entropy_models/entropy_models.py (EntropyModel :70-290,
EntropyBottleneck :293-522, GaussianConditional :525-666): same constructor arguments, parameter / buffer
names (so reference checkpoints load), methods, return shapes and ValueErrors.  What differs is where
the work happens:

  quantize / dequantize / build_indexes / likelihood   -> csrc/entropy.cu (one fused launch each)
  compress / decompress                                -> csrc/rans.cu, all images of the batch in one
                                                          launch, tables resident on the device
  update() (one-time table construction)               -> pmf in fp32 on the host exactly as the reference
                                                          does on CPU, then the C-ABI pmf_to_quantized_cdf

Inference needs CUDA tensors (no CPU implementation).  The training-mode forward (additive uniform noise,
autograd) is plain PyTorch and is not part of the accelerated path.
"""
import ctypes as C
import math
from typing import Any, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from compressai import _native, ans
from compressai._CXX import pmf_to_quantized_cdf as _pmf_to_quantized_cdf
from compressai._native import NULL_VIEW, NativeError, check, lib, stream_ptr, view_nchw
from compressai.ops import LowerBound

# -Phi^-1(1e-9 / 2): the reference evaluates scipy.stats.norm.ppf (entropy_models.py:586,600); the
# constant is used when scipy is unavailable and is checked against scipy in tests/.
_GAUSS_TAIL_MULTIPLIER = 6.109410204869


class _EntropyCoder:
    """Proxy to the coder implementation (reference :17-51); only "ans" exists in this package."""

    def __init__(self, method):
        if not isinstance(method, str):
            raise ValueError(f'Invalid method type "{type(method)}"')
        from compressai import available_entropy_coders

        if method not in available_entropy_coders():
            methods = ", ".join(available_entropy_coders())
            raise ValueError(f'Unknown entropy coder "{method}" (available: {methods})')
        self.name = method
        self._encoder = ans.RansEncoder()
        self._decoder = ans.RansDecoder()

    def encode_with_indexes(self, *args, **kwargs):
        return self._encoder.encode_with_indexes(*args, **kwargs)

    def decode_with_indexes(self, *args, **kwargs):
        return self._decoder.decode_with_indexes(*args, **kwargs)


def default_entropy_coder():
    from compressai import get_entropy_coder

    return get_entropy_coder()


def pmf_to_quantized_cdf(pmf: Tensor, precision: int = 16) -> Tensor:
    return torch.IntTensor(_pmf_to_quantized_cdf(pmf.tolist(), precision))


def _need_cuda(t, what):
    if not t.is_cuda:
        raise NativeError(f"{what}: tensor is on {t.device}; this package implements the entropy-model hot path on CUDA only")


def _bcp(t):
    """(B, C, P) of an N-d tensor with batch and channel leading."""
    return t.shape[0], t.shape[1], int(np.prod(t.shape[2:])) if t.dim() > 2 else 1


class EntropyModel(nn.Module):
    def __init__(self, likelihood_bound: float = 1e-9, entropy_coder: Optional[str] = None, entropy_coder_precision: int = 16):
        super().__init__()
        if entropy_coder is None:
            entropy_coder = default_entropy_coder()
        self.entropy_coder = _EntropyCoder(entropy_coder)
        self.entropy_coder_precision = int(entropy_coder_precision)
        self.likelihood_bound = float(likelihood_bound)
        self.use_likelihood_bound = likelihood_bound > 0
        if self.use_likelihood_bound:
            self.likelihood_lower_bound = LowerBound(likelihood_bound)
        self.register_buffer("_offset", torch.IntTensor())
        self.register_buffer("_quantized_cdf", torch.IntTensor())
        self.register_buffer("_cdf_length", torch.IntTensor())
        self._dev_tables = None
        self._dev_tables_key = None

    def __getstate__(self):
        d = self.__dict__.copy()
        d["entropy_coder"] = self.entropy_coder.name
        d["_dev_tables"] = None
        d["_dev_tables_key"] = None
        return d

    def __setstate__(self, state):
        self.__dict__ = state
        self.entropy_coder = _EntropyCoder(self.__dict__.pop("entropy_coder"))

    @property
    def offset(self):
        return self._offset

    @property
    def quantized_cdf(self):
        return self._quantized_cdf

    @property
    def cdf_length(self):
        return self._cdf_length

    def forward(self, *args: Any) -> Any:
        raise NotImplementedError()

    # -- device-resident tables -----------------------------------------------------------------
    def device_tables(self) -> ans.Tables:
        """icm_tables built from the current buffers (rebuilt when update()/load_state_dict change them)."""
        self._check_cdf_size()
        self._check_cdf_length()
        self._check_offsets_size()
        q = self._quantized_cdf
        key = (q.data_ptr(), q._version, self._cdf_length._version, self._offset._version, tuple(q.shape), torch.cuda.current_device())
        if self._dev_tables is None or self._dev_tables_key != key:
            self._dev_tables = ans.Tables(q, self._cdf_length, self._offset)
            self._dev_tables_key = key
        return self._dev_tables

    # -- E1 / E2 -----------------------------------------------------------------------------------
    def quantize(self, inputs: Tensor, mode: str, means: Optional[Tensor] = None) -> Tensor:
        if mode not in ("noise", "dequantize", "symbols"):
            raise ValueError(f'Invalid quantization mode: "{mode}"')
        if mode == "noise":  # training only
            return inputs + torch.empty_like(inputs).uniform_(-0.5, 0.5)
        _need_cuda(inputs, "quantize")
        x = inputs.detach().float().contiguous()
        lead = x.shape if x.dim() >= 2 else (1, 1) + tuple(x.shape)
        xv = x.reshape(lead)
        B, Cc, P = _bcp(xv)
        mv = NULL_VIEW
        if means is not None:
            m = means.detach().float().expand_as(x).contiguous().reshape(lead)
            mv = view_nchw(m)
        sym = torch.empty(lead, dtype=torch.int32, device=x.device)
        out = torch.empty_like(xv) if mode == "dequantize" else None
        check(lib().icm_gc_quantize_index(view_nchw(xv), mv, NULL_VIEW, B, Cc, P, None, 0, 0.0, sym.data_ptr(), None, Cc * P, 0,
                                          view_nchw(out), NULL_VIEW, NULL_VIEW, stream_ptr()), "icm_gc_quantize_index")
        return out.reshape(inputs.shape) if mode == "dequantize" else sym.reshape(inputs.shape)

    def _quantize(self, inputs, mode, means=None):
        return self.quantize(inputs, mode, means)

    @staticmethod
    def dequantize(inputs: Tensor, means: Optional[Tensor] = None) -> Tensor:
        if means is None:
            return inputs.float()
        _need_cuda(inputs, "dequantize")
        s = inputs.to(torch.int32).contiguous()
        lead = s.shape if s.dim() >= 2 else (1, 1) + tuple(s.shape)
        m = means.detach().float().expand_as(s).contiguous().reshape(lead)
        B, Cc, P = _bcp(m)
        out = torch.empty_like(m)
        check(lib().icm_gc_dequantize(s.data_ptr(), Cc * P, 0, view_nchw(m), B, Cc, P, view_nchw(out), NULL_VIEW, NULL_VIEW, stream_ptr()),
              "icm_gc_dequantize")
        return out.reshape(inputs.shape).type_as(means)

    @classmethod
    def _dequantize(cls, inputs, means=None):
        return cls.dequantize(inputs, means)

    # -- E8 helper -----------------------------------------------------------------------------------
    def _pmf_to_cdf(self, pmf, tail_mass, pmf_length, max_length):
        cdf = torch.zeros((len(pmf_length), max_length + 2), dtype=torch.int32)
        pmf, tail_mass = pmf.detach().cpu(), tail_mass.detach().cpu()
        for i in range(len(pmf_length)):
            prob = torch.cat((pmf[i, : int(pmf_length[i])], tail_mass[i]), dim=0)
            row = pmf_to_quantized_cdf(prob, self.entropy_coder_precision)
            cdf[i, : row.size(0)] = row
        return cdf

    def _check_cdf_size(self):
        if self._quantized_cdf.numel() == 0:
            raise ValueError("Uninitialized CDFs. Run update() first")
        if self._quantized_cdf.dim() != 2:
            raise ValueError(f"Invalid CDF size {self._quantized_cdf.size()}")

    def _check_offsets_size(self):
        if self._offset.numel() == 0:
            raise ValueError("Uninitialized offsets. Run update() first")
        if self._offset.dim() != 1:
            raise ValueError(f"Invalid offsets size {self._offset.size()}")

    def _check_cdf_length(self):
        if self._cdf_length.numel() == 0:
            raise ValueError("Uninitialized CDF lengths. Run update() first")
        if self._cdf_length.dim() != 1:
            raise ValueError(f"Invalid offsets size {self._cdf_length.size()}")

    # -- E7 ------------------------------------------------------------------------------------------
    def compress(self, inputs, indexes, means=None, flag=1):
        """inputs/indexes [B, C, ...] -> list of B byte strings (one rANS stream per image)."""
        if inputs.dim() < 2:
            raise ValueError("Invalid `inputs` size. Expected a tensor with at least 2 dimensions.")
        if inputs.size() != indexes.size():
            raise ValueError("`inputs` and `indexes` should have the same size.")
        tables = self.device_tables()
        symbols = self.quantize(inputs, "symbols", means)
        B = symbols.size(0)
        idx = indexes.to(device=symbols.device, dtype=torch.int32)
        return ans.encode_streams(tables, symbols.reshape(B, -1), idx.reshape(B, -1))

    def decompress(self, strings, indexes, means=None, flag=1):
        if not isinstance(strings, (tuple, list)):
            raise ValueError("Invalid `strings` parameter type.")
        if not len(strings) == indexes.size(0):
            raise ValueError("Invalid strings or indexes parameters")
        if indexes.dim() < 2:
            raise ValueError("Invalid `indexes` size. Expected a tensor with at least 2 dimensions.")
        tables = self.device_tables()
        if means is not None:
            if means.size()[:2] != indexes.size()[:2]:
                raise ValueError("Invalid means or indexes parameters")
            if means.size() != indexes.size():
                for i in range(2, indexes.dim()):
                    if means.size(i) != 1:
                        raise ValueError("Invalid means parameters")
        dev = self._quantized_cdf.device if self._quantized_cdf.is_cuda else torch.device("cuda", torch.cuda.current_device())
        idx = indexes.to(device=dev, dtype=torch.int32).contiguous()
        B = idx.size(0)
        dec = ans.StreamDecoder(B)
        dec.set_streams(strings)
        sym = dec.decode_step(tables, idx.reshape(B, -1)).reshape(idx.shape)
        dec.check_status()
        if means is not None:
            means = means.to(dev).expand_as(sym)
        return self.dequantize(sym, means)


class EntropyBottleneck(EntropyModel):
    """Fully-factorised prior (Ballé et al. 2018); reference :293-522."""

    _offset: Tensor

    def __init__(self, channels: int, *args: Any, tail_mass: float = 1e-9, init_scale: float = 10,
                 filters: Tuple[int, ...] = (3, 3, 3, 3), **kwargs: Any):
        super().__init__(*args, **kwargs)
        self.channels = int(channels)
        self.filters = tuple(int(f) for f in filters)
        self.init_scale = float(init_scale)
        self.tail_mass = float(tail_mass)
        dims = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        for i in range(len(self.filters) + 1):
            init = np.log(np.expm1(1 / scale / dims[i + 1]))
            self.register_parameter(f"_matrix{i:d}", nn.Parameter(torch.full((self.channels, dims[i + 1], dims[i]), float(init))))
            self.register_parameter(f"_bias{i:d}", nn.Parameter(torch.empty(self.channels, dims[i + 1], 1).uniform_(-0.5, 0.5)))
            if i < len(self.filters):
                self.register_parameter(f"_factor{i:d}", nn.Parameter(torch.zeros(self.channels, dims[i + 1], 1)))
        self.quantiles = nn.Parameter(torch.tensor([-self.init_scale, 0.0, self.init_scale]).repeat(self.channels, 1, 1))
        target = np.log(2 / self.tail_mass - 1)
        self.register_buffer("target", torch.Tensor([-target, 0, target]))
        self._packed = None
        self._packed_key = None

    def _get_medians(self) -> Tensor:
        return self.quantiles[:, :, 1:2]

    # one-time table construction (E8), computed on the host in fp32 like the reference's CPU path
    def update(self, force: bool = False) -> bool:
        if self._offset.numel() > 0 and not force:
            return False
        dev = self.quantiles.device
        p = {k: v.detach().float().cpu() for k, v in self.named_parameters(recurse=False)}
        q = p["quantiles"]
        medians = q[:, 0, 1]
        minima = torch.clamp(torch.ceil(medians - q[:, 0, 0]).int(), min=0)
        maxima = torch.clamp(torch.ceil(q[:, 0, 2] - medians).int(), min=0)
        pmf_start = medians - minima
        pmf_length = maxima + minima + 1
        max_length = int(pmf_length.max())
        samples = torch.arange(max_length)[None, :] + pmf_start[:, None, None]
        lower = self._logits_cumulative(samples - 0.5, True, p)
        upper = self._logits_cumulative(samples + 0.5, True, p)
        sign = -torch.sign(lower + upper)
        pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))[:, 0, :]
        tail_mass = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
        self._quantized_cdf = self._pmf_to_cdf(pmf, tail_mass, pmf_length, max_length).to(dev)
        self._offset = (-minima).to(dev)
        self._cdf_length = (pmf_length + 2).to(dev)
        return True

    def loss(self) -> Tensor:
        logits = self._logits_cumulative(self.quantiles, stop_gradient=True)
        return torch.abs(logits - self.target).sum()

    def _logits_cumulative(self, inputs: Tensor, stop_gradient: bool, params=None) -> Tensor:
        """PyTorch evaluation of the cumulative logits (reference :400-419); used by update(), loss() and
        the training-mode forward.  The eval forward uses the fused kernel instead."""
        get = (lambda n: params[n]) if params is not None else (lambda n: getattr(self, n))
        h = inputs
        for i in range(len(self.filters) + 1):
            m, b = get(f"_matrix{i:d}"), get(f"_bias{i:d}")
            if stop_gradient:
                m, b = m.detach(), b.detach()
            h = torch.matmul(F.softplus(m), h) + b
            if i < len(self.filters):
                f = get(f"_factor{i:d}")
                if stop_gradient:
                    f = f.detach()
                h = h + torch.tanh(f) * torch.tanh(h)
        return h

    def _likelihood(self, inputs: Tensor) -> Tensor:
        lower = self._logits_cumulative(inputs - 0.5, stop_gradient=False)
        upper = self._logits_cumulative(inputs + 0.5, stop_gradient=False)
        sign = -torch.sign(lower + upper).detach()
        return torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))

    def packed_params(self) -> Tensor:
        """[C, 59] fp32 on the device: the per-channel record icm_eb_process expects (include/icm_b200.h)."""
        if self.filters != (3, 3, 3, 3):
            raise NativeError("the CUDA EntropyBottleneck kernel is built for filters=(3,3,3,3) (the reference's only configuration)")
        ps = [getattr(self, f"_matrix{i}") for i in range(5)] + [getattr(self, f"_bias{i}") for i in range(5)] + \
             [getattr(self, f"_factor{i}") for i in range(4)] + [self.quantiles]
        key = tuple((p.data_ptr(), p._version) for p in ps)
        if self._packed is None or self._packed_key != key:
            Cn = self.channels
            parts = []
            for i in range(5):
                parts.append(getattr(self, f"_matrix{i}").detach().reshape(Cn, -1))
                parts.append(getattr(self, f"_bias{i}").detach().reshape(Cn, -1))
                if i < 4:
                    parts.append(getattr(self, f"_factor{i}").detach().reshape(Cn, -1))
            parts.append(self.quantiles.detach()[:, 0, 1:2])
            self._packed = torch.cat(parts, 1).float().contiguous()
            assert self._packed.shape[1] == _native.EB_PARAMS
            self._packed_key = key
            if self._packed.is_cuda:
                torch.cuda.current_stream().synchronize()  # ready before another stream may read it
        return self._packed

    def forward(self, x: Tensor, training: Optional[bool] = None) -> Tuple[Tensor, Tensor]:
        if training is None:
            training = self.training
        if training or (torch.is_grad_enabled() and x.requires_grad):
            return self._forward_torch(x, training)
        _need_cuda(x, "EntropyBottleneck.forward")
        xv = x.detach().float().contiguous()
        B, Cc, P = _bcp(xv)
        z_hat, lik = torch.empty_like(xv), torch.empty_like(xv)
        check(lib().icm_eb_process(1, view_nchw(xv), B, Cc, P, self.packed_params().data_ptr(),
                                   self.likelihood_bound if self.use_likelihood_bound else 0.0, None, None,
                                   view_nchw(z_hat), NULL_VIEW, view_nchw(lik), stream_ptr()), "icm_eb_process")
        return z_hat, lik

    def _forward_torch(self, x, training):
        perm = list(range(x.dim()))
        perm[0], perm[1] = perm[1], perm[0]
        v = x.permute(*perm).contiguous()
        shape = v.size()
        v = v.reshape(v.size(0), 1, -1)
        med = self._get_medians()
        out = v + torch.empty_like(v).uniform_(-0.5, 0.5) if training else torch.round(v - med) + med
        lik = self._likelihood(out)
        if self.use_likelihood_bound:
            lik = self.likelihood_lower_bound(lik)
        back = lambda t: t.reshape(shape).permute(*perm).contiguous()
        return back(out), back(lik)

    @staticmethod
    def _build_indexes(size):
        N, Cn = size[0], size[1]
        view = [1] * len(size)
        view[1] = -1
        return torch.arange(Cn).view(*view).int().repeat(N, 1, *size[2:])

    @staticmethod
    def _extend_ndims(tensor, n):
        return tensor.reshape(-1, *([1] * n)) if n > 0 else tensor.reshape(-1)

    def compress(self, x):
        """z [B, C, ...] -> list of B strings (reference :508-515); symbols/indexes come out of one launch."""
        _need_cuda(x, "EntropyBottleneck.compress")
        tables = self.device_tables()
        xv = x.detach().float().contiguous()
        B, Cc, P = _bcp(xv)
        sym = torch.empty((B, Cc * P), dtype=torch.int32, device=x.device)
        idx = torch.empty_like(sym)
        check(lib().icm_eb_process(0, view_nchw(xv), B, Cc, P, self.packed_params().data_ptr(), 0.0, sym.data_ptr(), idx.data_ptr(),
                                   NULL_VIEW, NULL_VIEW, NULL_VIEW, stream_ptr()), "icm_eb_process")
        return ans.encode_streams(tables, sym, idx)

    def decompress(self, strings, size):
        output_size = (len(strings), self._quantized_cdf.size(0), *size)
        indexes = self._build_indexes(output_size).to(self._quantized_cdf.device)
        medians = self._extend_ndims(self._get_medians().detach(), len(size))
        medians = medians.expand(len(strings), *([-1] * (len(size) + 1)))
        return super().decompress(strings, indexes, medians, 0)


class GaussianConditional(EntropyModel):
    """Gaussian conditional with a discrete scale table; reference :525-666."""

    def __init__(self, scale_table, *args: Any, scale_bound: float = 0.11, tail_mass: float = 1e-9, **kwargs: Any):
        super().__init__(*args, **kwargs)
        if not isinstance(scale_table, (type(None), list, tuple)):
            raise ValueError(f'Invalid type for scale_table "{type(scale_table)}"')
        if isinstance(scale_table, (list, tuple)) and len(scale_table) < 1:
            raise ValueError(f'Invalid scale_table length "{len(scale_table)}"')
        if scale_table and (scale_table != sorted(scale_table) or any(s <= 0 for s in scale_table)):
            raise ValueError(f'Invalid scale_table "({scale_table})"')
        self.tail_mass = float(tail_mass)
        if scale_bound is None and scale_table:
            scale_bound = scale_table[0]
        if scale_bound <= 0:
            raise ValueError("Invalid parameters")
        self.lower_bound_scale = LowerBound(scale_bound)
        self.register_buffer("scale_table", self._prepare_scale_table(scale_table) if scale_table else torch.Tensor())
        self.register_buffer("scale_bound", torch.Tensor([float(scale_bound)]) if scale_bound is not None else None)
        self._scale_bound_f = float(scale_bound)

    @staticmethod
    def _prepare_scale_table(scale_table):
        return torch.Tensor(tuple(float(s) for s in scale_table))

    def _standardized_cumulative(self, inputs: Tensor) -> Tensor:
        return 0.5 * torch.erfc(float(-(2 ** -0.5)) * inputs)

    @staticmethod
    def _standardized_quantile(quantile):
        try:
            import scipy.stats

            return float(scipy.stats.norm.ppf(quantile))
        except ImportError:  # pragma: no cover
            if abs(quantile - 0.5e-9) > 1e-24:
                raise
            return -_GAUSS_TAIL_MULTIPLIER

    def update_scale_table(self, scale_table, force=False):
        if self._offset.numel() > 0 and not force:
            return False
        device = self.scale_table.device
        self.scale_table = self._prepare_scale_table(scale_table).to(device)
        self.update()
        return True

    def update(self):
        """Table construction (reference :599-624) in fp32 on the host, bit-identical to the reference's CPU result."""
        dev = self.scale_table.device
        table = self.scale_table.detach().float().cpu()
        multiplier = -self._standardized_quantile(self.tail_mass / 2)
        pmf_center = torch.ceil(table * multiplier).int()
        pmf_length = 2 * pmf_center + 1
        max_length = int(torch.max(pmf_length))
        samples = torch.abs(torch.arange(max_length).int() - pmf_center[:, None]).float()
        scale = table.unsqueeze(1)
        upper = self._standardized_cumulative((0.5 - samples) / scale)
        lower = self._standardized_cumulative((-0.5 - samples) / scale)
        pmf = upper - lower
        tail_mass = 2 * lower[:, :1]
        self._quantized_cdf = self._pmf_to_cdf(pmf, tail_mass, pmf_length, max_length).to(dev)
        self._offset = (-pmf_center).to(dev)
        self._cdf_length = (pmf_length + 2).to(dev)

    def _likelihood(self, inputs: Tensor, scales: Tensor, means: Optional[Tensor] = None) -> Tensor:
        """PyTorch form (reference :626-643) for the training path."""
        values = inputs if means is None else inputs - means
        scales = self.lower_bound_scale(scales)
        values = torch.abs(values)
        upper = self._standardized_cumulative((0.5 - values) / scales)
        lower = self._standardized_cumulative((-0.5 - values) / scales)
        return upper - lower

    def forward(self, inputs: Tensor, scales: Tensor, means: Optional[Tensor] = None, training: Optional[bool] = None):
        if training is None:
            training = self.training
        if training or (torch.is_grad_enabled() and (inputs.requires_grad or scales.requires_grad)):
            outputs = self.quantize(inputs, "noise", means) if training else _round_keep(inputs, means)
            lik = self._likelihood(outputs, scales, means)
            if self.use_likelihood_bound:
                lik = self.likelihood_lower_bound(lik)
            return outputs, lik
        _need_cuda(inputs, "GaussianConditional.forward")
        x = inputs.detach().float().contiguous()
        s = scales.detach().float().expand_as(x).contiguous()
        B, Cc, P = _bcp(x)
        mv = NULL_VIEW
        if means is not None:
            m = means.detach().float().expand_as(x).contiguous()
            mv = view_nchw(m)
        y_hat, lik = torch.empty_like(x), torch.empty_like(x)
        check(lib().icm_gc_likelihood(view_nchw(x), mv, view_nchw(s), B, Cc, P, self._scale_bound_f,
                                      self.likelihood_bound if self.use_likelihood_bound else 0.0,
                                      view_nchw(y_hat), view_nchw(lik), NULL_VIEW, NULL_VIEW, stream_ptr()), "icm_gc_likelihood")
        return y_hat, lik

    def scale_table_device(self, device):
        t = self.scale_table
        if t.device != device:
            t = t.to(device)
        return t.float().contiguous()

    def build_indexes(self, scales: Tensor) -> Tensor:
        """E3: one launch instead of len(scale_table)-1 compare/subtract passes (reference :661-666)."""
        _need_cuda(scales, "GaussianConditional.build_indexes")
        s = scales.detach().float().contiguous()
        lead = s.shape if s.dim() >= 2 else (1, 1) + tuple(s.shape)
        sv = s.reshape(lead)
        B, Cc, P = _bcp(sv)
        table = self.scale_table_device(s.device)
        idx = torch.empty(lead, dtype=torch.int32, device=s.device)
        check(lib().icm_gc_build_indexes(view_nchw(sv), B, Cc, P, table.data_ptr(), table.numel(), self._scale_bound_f,
                                         idx.data_ptr(), Cc * P, 0, stream_ptr()), "icm_gc_build_indexes")
        return idx.reshape(scales.shape)


def _round_keep(inputs, means):
    if means is None:
        return torch.round(inputs)
    return torch.round(inputs - means) + means
