from .Model import Model
from .TransE import TransE
from .TransH import TransH
from .TransD import TransD

__all__ = ["Model", "TransE", "TransH", "TransD"]
