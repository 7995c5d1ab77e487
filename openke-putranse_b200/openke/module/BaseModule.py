"""Checkpoint / parameter helpers shared by models, losses and strategies.

Mirrors the interface of the reference's ``openke/module/BaseModule.py:8-54`` (same method names,
the same two frozen constants so that ``state_dict()`` keys match the shipped checkpoints:
``zero_const``, ``pi_const``), with ``map_location`` added so CUDA-saved checkpoints load anywhere.
"""
import json

import torch
import torch.nn as nn


class BaseModule(nn.Module):
    def __init__(self):
        super().__init__()
        self.zero_const = nn.Parameter(torch.tensor([0.0]), requires_grad=False)
        self.pi_const = nn.Parameter(torch.tensor([3.14159265358979323846]), requires_grad=False)

    # -- torch checkpoints
    def load_checkpoint(self, path, map_location=None):
        self.load_state_dict(torch.load(path, map_location=map_location or self.zero_const.device))
        self.eval()

    def save_checkpoint(self, path):
        torch.save(self.state_dict(), path)

    # -- JSON parameters
    def get_parameters(self, mode="numpy", param_dict=None):
        state = self.state_dict()
        names = state.keys() if param_dict is None else param_dict
        conv = {"numpy": lambda v: v.cpu().numpy(), "list": lambda v: v.cpu().numpy().tolist()}.get(mode, lambda v: v)
        return {name: conv(state[name]) for name in names}

    def set_parameters(self, parameters):
        self.load_state_dict({k: torch.as_tensor(v, dtype=torch.float32) for k, v in parameters.items()}, strict=False)
        self.eval()

    def save_parameters(self, path):
        with open(path, "w") as f:
            f.write(json.dumps(self.get_parameters("list")))

    def load_parameters(self, path):
        with open(path, "r") as f:
            self.set_parameters(json.loads(f.read()))
