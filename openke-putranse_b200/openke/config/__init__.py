from .Trainer import Trainer
from .Tester import Tester
from .Validator import Validator
from .Parallel_Universe_Config import Parallel_Universe_Config

__all__ = ["Trainer", "Tester", "Validator", "Parallel_Universe_Config"]
