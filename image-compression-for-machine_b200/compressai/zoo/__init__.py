"""Model registry (reference: /root/reference/compressai/zoo/__init__.py:21-43).  Only the architectures on
the accelerated hot path are registered; the reference's research variants (stf2..stf14, czigzag, ICM
detectors) are out of scope."""
from compressai.models import WACNN, WACNN2, SymmetricalTransFormer

from .pretrained import load_pretrained
from .pretrained import load_pretrained as load_state_dict  # the reference's alias (zoo/__init__.py:21)

models = {"stf": SymmetricalTransFormer, "cnn": WACNN, "cnn2": WACNN2}

__all__ = ["models", "load_pretrained", "load_state_dict"]
