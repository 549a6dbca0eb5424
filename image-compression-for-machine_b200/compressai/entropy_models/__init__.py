from .entropy_models import EntropyBottleneck, EntropyModel, GaussianConditional, pmf_to_quantized_cdf

__all__ = ["EntropyModel", "EntropyBottleneck", "GaussianConditional", "pmf_to_quantized_cdf"]
