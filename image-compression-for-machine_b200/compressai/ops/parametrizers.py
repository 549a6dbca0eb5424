"""`NonNegativeParametrizer` (reference: /root/reference/compressai/ops/parametrizers.py:23-49).

Stored value v -> max(v, sqrt(minimum + pedestal))**2 - pedestal with pedestal = reparam_offset**2.
The GDN kernels consume the already re-parametrised beta/gamma (computed once per weight update)."""
import torch
import torch.nn as nn

from .bound_ops import LowerBound


class NonNegativeParametrizer(nn.Module):
    def __init__(self, minimum=0, reparam_offset=2 ** -18):
        super().__init__()
        self.minimum = float(minimum)
        self.reparam_offset = float(reparam_offset)
        pedestal = self.reparam_offset ** 2
        self.register_buffer("pedestal", torch.Tensor([pedestal]))
        self.lower_bound = LowerBound((self.minimum + pedestal) ** 0.5)

    def init(self, x):
        return torch.sqrt(torch.max(x + self.pedestal, self.pedestal))

    def forward(self, x):
        return self.lower_bound(x) ** 2 - self.pedestal
