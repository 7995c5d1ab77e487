"""Filtered hits@10 on the validation split, for early stopping of single-space models (interface of
the reference's ``openke/config/Validator.py:20-53``)."""
import numpy as np
import torch

from .Tester import Tester


class Validator(Tester):
    def __init__(self, model=None, data_loader=None):
        super().__init__(model=model, data_loader=data_loader, use_gpu=torch.cuda.is_available())
        self.valid_dataloader = data_loader
        self.early_stopping_patience = 10
        self.bad_counts = 0
        self.best_hit10 = 0

    def valid(self):
        ranks = self.rank_all(loader=self.valid_dataloader)
        n = np.float32(ranks.shape[0])
        # reference Valid.h:242-257: per-side float32 fractions, then their mean
        l10 = np.float32((ranks[:, 1] < 10).sum()) / n
        r10 = np.float32((ranks[:, 3] < 10).sum()) / n
        return float((l10 + r10) / np.float32(2))

    def valid_one_step(self, data):
        return self.model.predict(data)
