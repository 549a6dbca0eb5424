"""CompressionModel base (reference: /root/reference/compressai/models/base.py:6-70)."""
import torch.nn as nn

from compressai.entropy_models import EntropyBottleneck

from .utils import update_registered_buffers


class CompressionModel(nn.Module):
    def __init__(self, init_weights=True):
        super().__init__()
        # The reference calls _initialize_weights() here, before any sub-module exists, so it is a no-op
        # there too (SURVEY.md §8d "Weights"): the effective init is PyTorch's per-layer default.

    def aux_loss(self):
        return sum(m.loss() for m in self.modules() if isinstance(m, EntropyBottleneck))

    def forward(self, *args):
        raise NotImplementedError()

    def update(self, force=False):
        updated = False
        for m in self.children():
            if isinstance(m, EntropyBottleneck):
                updated |= m.update(force=force)
        self._native_cache_clear()
        return updated

    def _native_cache_clear(self):
        eng = getattr(self, "_engine", None)
        if eng is not None:
            eng.invalidate()

    def load_state_dict(self, state_dict, strict=False):
        update_registered_buffers(self.entropy_bottleneck, "entropy_bottleneck", ["_quantized_cdf", "_offset", "_cdf_length"], state_dict)
        out = super().load_state_dict(state_dict, strict=strict)
        self._native_cache_clear()
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._native_cache_clear()
        return out
