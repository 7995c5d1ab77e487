"""Pairs every positive score with its k negatives and applies the loss (reference
openke/module/strategy/NegativeSampling.py:5-33): scores arrive as [B positives | B neg#1 | ...].
``Trainer`` does not call ``forward`` — it hands model, loss margin and batch size to the fused CUDA
step — but the object keeps the reference's attributes so existing scripts run unchanged."""
from .Strategy import Strategy


class NegativeSampling(Strategy):
    def __init__(self, model=None, loss=None, batch_size=256, regul_rate=0.0, l3_regul_rate=0.0):
        super().__init__()
        if regul_rate != 0.0 or l3_regul_rate != 0.0:
            raise NotImplementedError("regularisation terms are not part of the fused PuTransE train step")
        self.model, self.loss, self.batch_size = model, loss, batch_size
        self.regul_rate, self.l3_regul_rate = regul_rate, l3_regul_rate

    def _get_positive_score(self, score):
        return score[:self.batch_size].view(-1, self.batch_size).permute(1, 0)

    def _get_negative_score(self, score):
        return score[self.batch_size:].view(-1, self.batch_size).permute(1, 0)

    def forward(self, data):
        score = self.model(data)
        return self.loss(self._get_positive_score(score), self._get_negative_score(score))
