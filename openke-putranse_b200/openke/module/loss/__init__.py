from .Loss import Loss
from .MarginLoss import MarginLoss

__all__ = ["Loss", "MarginLoss"]
