"""Small torch modules the LSTEP drop-in needs with the reference's parameter names, so that
reference checkpoints load unchanged (models/modules.py:7-68). These are plain PyTorch: the
link predictor and the time encoder's parameters are outside the CUDA hot path (the time
encoding itself is fused into the aggregation kernels)."""
import numpy as np
import torch
import torch.nn as nn


class TimeEncoder(nn.Module):
    """cos(t * w + b) with w_j = 10^(-9 j / (dim - 1)) in fp32 and b = 0 (models/modules.py:20-21)."""

    def __init__(self, time_dim: int, parameter_requires_grad: bool = True):
        super().__init__()
        self.time_dim = time_dim
        self.w = nn.Linear(1, time_dim)
        freqs = 1 / 10 ** np.linspace(0, 9, time_dim, dtype=np.float32)
        self.w.weight = nn.Parameter(torch.from_numpy(freqs).reshape(time_dim, 1))
        self.w.bias = nn.Parameter(torch.zeros(time_dim))
        if not parameter_requires_grad:
            self.w.weight.requires_grad_(False)
            self.w.bias.requires_grad_(False)

    def forward(self, timestamps: torch.Tensor) -> torch.Tensor:
        # (batch, seq) -> (batch, seq, time_dim)
        return torch.cos(self.w(timestamps.unsqueeze(2)))


class MergeLayer(nn.Module):
    """Link predictor head: fc2(relu(fc1([x1 || x2]))) (models/modules.py:42-68)."""

    def __init__(self, input_dim1: int, input_dim2: int, hidden_dim: int, output_dim: int):
        super().__init__()
        self.fc1 = nn.Linear(input_dim1 + input_dim2, hidden_dim)
        self.fc2 = nn.Linear(hidden_dim, output_dim)
        self.act = nn.ReLU()

    def forward(self, input_1: torch.Tensor, input_2: torch.Tensor) -> torch.Tensor:
        return self.fc2(self.act(self.fc1(torch.cat([input_1, input_2], dim=1))))
