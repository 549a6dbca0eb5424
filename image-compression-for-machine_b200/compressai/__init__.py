"""B200-native drop-in for the `compressai` package of stm233/image-compression-for-machine.

Same import surface as the reference for its inference hot path
(/root/reference/compressai/__init__.py:15-62): `compressai.ans`, `compressai._CXX`,
`compressai.entropy_models`, `compressai.layers`, `compressai.ops`, `compressai.models`, `compressai.zoo`
and the entropy-coder registry below; underneath, every operator calls hand-written sm_100a CUDA kernels
through the C ABI in include/icm_b200.h.  There is no CPU path: tensors must live on a CUDA device.
"""
__version__ = "1.1.6.dev0+b200"

_entropy_coder = "ans"
_available_entropy_coders = [_entropy_coder]


def set_entropy_coder(entropy_coder):
    """Select the default entropy coder (only "ans" exists here; reference: __init__.py:33-48)."""
    global _entropy_coder
    if entropy_coder not in _available_entropy_coders:
        raise ValueError(
            f'Invalid entropy coder "{entropy_coder}", choose from({", ".join(_available_entropy_coders)}).'
        )
    _entropy_coder = entropy_coder


def get_entropy_coder():
    return _entropy_coder


def available_entropy_coders():
    return _available_entropy_coders


from compressai import entropy_models, layers, models, ops  # noqa: E402,F401
