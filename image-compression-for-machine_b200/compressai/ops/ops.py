"""`ste_round` (reference: /root/reference/compressai/ops/ops.py:20-34).

Training-time helper (straight-through rounding).  The inference kernels fuse the rounding into the
entropy-stage kernels (csrc/entropy.cu); this function exists for API parity and for the training row."""
import torch


class _RoundSTE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return torch.round(x)

    @staticmethod
    def backward(ctx, g):
        return g


def ste_round(x):
    """round(x) in the forward pass, identity gradient in the backward pass."""
    return _RoundSTE.apply(x)
