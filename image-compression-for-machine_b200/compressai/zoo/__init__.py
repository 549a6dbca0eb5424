"""Model registry (reference: /root/reference/compressai/zoo/__init__.py:23-43).  Only the architectures on
the accelerated hot path are registered; the reference's research variants (stf2..stf14, czigzag, ICM
detectors) are out of scope."""
from compressai.models import SymmetricalTransFormer

from .pretrained import load_pretrained

models = {"stf": SymmetricalTransFormer}
try:
    from compressai.models.cnn import WACNN
    models["cnn"] = WACNN
except ImportError:  # WACNN lands in a later milestone
    pass

__all__ = ["models", "load_pretrained"]
