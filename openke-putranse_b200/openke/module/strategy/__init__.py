from .Strategy import Strategy
from .NegativeSampling import NegativeSampling

__all__ = ["Strategy", "NegativeSampling"]
