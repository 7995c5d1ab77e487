"""B200-native drop-in for the hot path of OpenKE-PuTransE.

Same import surface as the reference package (``from openke.config import Trainer, Tester,
Parallel_Universe_Config``; ``from openke.data import TrainDataLoader, TestDataLoader``;
``from openke.module.model import TransE, TransH, TransD`` ...), but the training step, the
batched universe training and the link-prediction ranking run in hand-written sm_100a CUDA kernels
behind the C-ABI of ``release/libputranse.so`` (see include/putranse.h and INTEGRATION.md).
"""
__version__ = "0.1"
