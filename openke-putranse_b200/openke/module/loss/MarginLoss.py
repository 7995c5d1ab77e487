"""Margin ranking loss  mean(max(p - n, -margin)) + margin  (reference
openke/module/loss/MarginLoss.py:8-28).  In training the loss and its gradient are computed inside
the fused CUDA step; ``forward`` here serves callers that hold score tensors.  The self-adversarial
weighting branch of the reference (``adv_temperature``) is not on the PuTransE hot path."""
import torch
import torch.nn as nn

from .Loss import Loss


class MarginLoss(Loss):
    def __init__(self, adv_temperature=None, margin=6.0):
        super().__init__()
        if adv_temperature is not None:
            raise NotImplementedError("adv_temperature is not supported by the fused train step")
        self.margin = nn.Parameter(torch.Tensor([margin]), requires_grad=False)
        self.adv_flag = False

    def forward(self, p_score, n_score):
        return torch.max(p_score - n_score, -self.margin).mean() + self.margin

    def predict(self, p_score, n_score):
        return self.forward(p_score, n_score).cpu().data.numpy()
