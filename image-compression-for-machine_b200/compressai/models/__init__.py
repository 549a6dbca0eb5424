from .base import CompressionModel
from .stf import SymmetricalTransFormer

__all__ = ["CompressionModel", "SymmetricalTransFormer"]
