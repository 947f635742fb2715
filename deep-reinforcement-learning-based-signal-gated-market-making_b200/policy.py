"""Policies and the (1,lambda) evolver with the reference's surface (models/model.py:5-76).

``TradingPolicy`` / ``AdversaryPolicy`` are ordinary ``nn.Module`` objects so that checkpoints
(``net.{0,2,4}.{weight,bias}`` / ``fc.{0,2}.{weight,bias}``) load unchanged and callers can keep
calling ``.forward`` in their per-bar loops.  The flat genome is ``parameters()`` order -- the
layout the CUDA kernels index directly (include/sgmm.h).
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn


def genome_len(hidden: int = 32) -> int:
    return hidden * hidden + 7 * hidden + 2


class _FlatMixin:
    def get_weights(self):
        return torch.cat([p.detach().reshape(-1) for p in self.parameters()]).clone()

    def set_weights(self, weights):
        w = torch.as_tensor(weights)
        at = 0
        for p in self.parameters():
            n = p.numel()
            p.data.copy_(w[at:at + n].reshape(p.shape).to(p.device))
            at += n


class TradingPolicy(_FlatMixin, nn.Module):
    """MLP state_dim -> hidden -> hidden -> action_dim with ReLU (model.py:7-15); orthogonal
    initialisation gain 0.9, biases 0.05 (model.py:18-21); inference only (model.py:22-26)."""

    def __init__(self, state_dim=3, action_dim=2, hidden_dim=32):
        super().__init__()
        layers = [nn.Linear(state_dim, hidden_dim), nn.ReLU(),
                  nn.Linear(hidden_dim, hidden_dim), nn.ReLU(),
                  nn.Linear(hidden_dim, action_dim)]
        self.net = nn.Sequential(*layers)
        for layer in layers:
            if isinstance(layer, nn.Linear):
                nn.init.orthogonal_(layer.weight, gain=0.9)
                nn.init.constant_(layer.bias, 0.05)
        self.eval()

    @torch.no_grad()
    def forward(self, x):
        return self.net(x)


class AdversaryPolicy(_FlatMixin, nn.Module):
    """3 -> 12 -> 2, ReLU then Tanh (model.py:40-50); default torch initialisation."""

    def __init__(self, input_size=3, hidden_size=12):
        super().__init__()
        self.fc = nn.Sequential(nn.Linear(input_size, hidden_size), nn.ReLU(),
                                nn.Linear(hidden_size, 2), nn.Tanh())

    def forward(self, x):
        return self.fc(x)


class NeuroEvolution:
    """(1,lambda)-ES bookkeeping (model.py:59-76): ``ask`` returns lambda children
    ``master + sigma*N(0,1)``, ``tell`` overwrites the master with the best child (no elitism,
    ``np.argmax`` = first maximum).  ``ask`` keeps the reference's host semantics (torch's global
    generator) for callers that want the list of tensors; :class:`engine.DRLEngine` does not call
    it -- it regenerates children on the device from a counter-based stream instead."""

    def __init__(self, population_size=50, sigma=0.05, hidden_dim=32):
        self.pop_size = population_size
        self.sigma = sigma
        self.master_policy = TradingPolicy(hidden_dim=hidden_dim)       # hidden_dim: models/model.py:7 (reference: 32)

    def ask(self):
        base = self.master_policy.get_weights()
        return [base + torch.randn_like(base) * self.sigma for _ in range(self.pop_size)]

    def tell(self, population_weights, fitness_scores):
        best = int(np.argmax(fitness_scores))
        self.master_policy.set_weights(population_weights[best])
        return fitness_scores[best]
