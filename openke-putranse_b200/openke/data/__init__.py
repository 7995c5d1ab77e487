from .TrainDataLoader import TrainDataLoader
from .TestDataLoader import TestDataLoader
from .IncrementalTrainDataLoader import IncrementalTrainDataLoader
from .IncrementalTestDataLoader import IncrementalTestDataLoader

__all__ = ["TrainDataLoader", "TestDataLoader", "IncrementalTrainDataLoader", "IncrementalTestDataLoader"]
