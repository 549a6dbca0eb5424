from .base import CompressionModel
from .cnn import WACNN, WACNN2
from .stf import SymmetricalTransFormer

__all__ = ["CompressionModel", "SymmetricalTransFormer", "WACNN", "WACNN2"]
