"""GDN parameter holder (reference: /root/reference/compressai/layers/gdn.py:26-75).

y_i = x_i * rsqrt(beta'_i + sum_j gamma'_ij x_j^2)  (inverse: * sqrt), with the non-negative re-parametrisation
beta' = max(beta, sqrt(beta_min + 2^-36))^2 - 2^-36, gamma' = max(gamma, 2^-18)^2 - 2^-36 (parametrizers.py:32-49).
The arithmetic runs as x^2 -> 1x1 tcgen05 GEMM with gamma' -> epilogue x * rsqrt(. + beta') (csrc/wacnn.cu, csrc/conv.cu)."""
import torch
import torch.nn as nn

from compressai.ops.parametrizers import NonNegativeParametrizer


class GDN(nn.Module):
    def __init__(self, in_channels, inverse=False, beta_min=1e-6, gamma_init=0.1):
        super().__init__()
        self.inverse = bool(inverse)
        self.beta_reparam = NonNegativeParametrizer(minimum=float(beta_min))
        self.beta = nn.Parameter(self.beta_reparam.init(torch.ones(in_channels)))
        self.gamma_reparam = NonNegativeParametrizer()
        self.gamma = nn.Parameter(self.gamma_reparam.init(float(gamma_init) * torch.eye(in_channels)))

    @torch.no_grad()
    def effective(self):
        """(gamma' [C_out, C_in], beta' [C]) as used by the kernels."""
        return self.gamma_reparam(self.gamma), self.beta_reparam(self.beta)
