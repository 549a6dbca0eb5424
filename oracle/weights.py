"""Deterministic, machine-independent synthetic weights for parity tests (TEST INFRASTRUCTURE).

No checkpoint ships with the reference and a 100 M-parameter state_dict cannot be committed, so the
golden fixtures in tests/golden/ are produced from weights that both sides can regenerate: every
floating-point parameter is drawn from a CPU torch.Generator seeded by (seed, crc32(name)), scaled
like PyTorch's default initialisers (the reference's effective init, SURVEY.md §8d "Weights").
`stress=True` applies the survey's rate-raising tweak so that many CDF tables and the bypass path
are exercised end to end (SURVEY.md §8d "Optional end-to-end stress weights").
"""
import math
import zlib

import torch

_SKIP = ("relative_position_index", "_offset", "_quantized_cdf", "_cdf_length", "target", "scale_table",
         "scale_bound", ".bound", "pedestal")


def _gen(seed, name):
    g = torch.Generator(device="cpu")
    g.manual_seed((int(seed) * 1_000_003 + zlib.crc32(name.encode())) % (2 ** 63))
    return g


def _uniform(shape, bound, g):
    return (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound


def seeded_state_dict(template, seed=0, stress=True):
    """template: a state_dict (names + shapes + dtypes) -> new state_dict with seeded values."""
    out = {}
    fan = {}
    for name, t in template.items():
        if name.endswith(".weight") and t.dim() >= 2:
            fan[name[: -len(".weight")]] = t[0].numel()
    for name, t in template.items():
        if any(s in name for s in _SKIP) or not t.is_floating_point():
            out[name] = t.clone()
            continue
        g = _gen(seed, name)
        base = name.rsplit(".", 1)[0]
        if "entropy_bottleneck" in name:
            if "_matrix" in name:
                v = t.clone() + _uniform(t.shape, 0.05, g)
            elif "_bias" in name:
                v = _uniform(t.shape, 0.5, g)
            elif "_factor" in name:
                v = _uniform(t.shape, 0.3, g)
            elif "quantiles" in name:
                v = t.clone()
                v[:, 0, 0] = -8.0 + _uniform((t.shape[0],), 3.0, g)
                v[:, 0, 1] = _uniform((t.shape[0],), 0.9, g)
                v[:, 0, 2] = 8.0 + _uniform((t.shape[0],), 3.0, g)
            else:
                v = t.clone()
        elif name.endswith("relative_position_bias_table"):
            v = _uniform(t.shape, 0.5, g)
        elif name.endswith(".beta"):  # GDN, reparametrised storage: sqrt(beta + pedestal)
            v = torch.sqrt(1.0 + _uniform(t.shape, 0.2, g).abs() + 2.0 ** -36)
        elif name.endswith(".gamma"):
            v = torch.sqrt(0.1 * torch.eye(t.shape[0]) + _uniform(t.shape, 0.01, g).abs() + 2.0 ** -36)
        elif t.dim() == 1 and ("norm" in name) and name.endswith(".weight"):
            v = 1.0 + _uniform(t.shape, 0.1, g)
        elif t.dim() == 1 and ("norm" in name) and name.endswith(".bias"):
            v = _uniform(t.shape, 0.1, g)
        elif t.dim() >= 2:
            v = _uniform(t.shape, 1.0 / math.sqrt(t[0].numel()), g)
        else:  # bias of a conv / linear
            v = _uniform(t.shape, 1.0 / math.sqrt(fan.get(base, t.numel())), g)
        out[name] = v.to(t.dtype)
    if stress:
        k = "layers.2.downsample.reduction.weight"
        if k in out:
            out[k] = out[k] * 8.0
        k = "g_a.7.weight"  # WACNN: the last analysis convolution sets the latent's magnitude
        if k in out:
            out[k] = out[k] * 48.0
        ramp = torch.exp(torch.linspace(math.log(0.05), math.log(30.0), 32))
        for name in out:
            if name.startswith("cc_scale_transforms.") and name.endswith(".8.bias"):
                out[name] = ramp.clone()
    return out


def seeded_image(shape, seed=0):
    g = torch.Generator(device="cpu")
    g.manual_seed(1234567 + int(seed))
    return torch.rand(shape, generator=g, dtype=torch.float32)
