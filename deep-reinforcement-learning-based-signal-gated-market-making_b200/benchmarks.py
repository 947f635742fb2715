"""Benchmark quoting rules with the reference's surface (Env/benchmarks.py:3-40) and their
device form: every rule here is a function of the inventory (and, for GLFT, of the bar's spread),
i.e. a ``[T, 5, 2]`` integer offset table walked by the same step core as the learned policy
(``engine.rollout_table``) -- one launch instead of the per-bar loop of main.py:99-132.
"""
from __future__ import annotations

import numpy as np


class FOICPolicy:
    """Fixed Offset with Inventory Constraints (Env/benchmarks.py:3-14)."""

    def __init__(self, offset_a=0, offset_b=0):
        self.offset_a = offset_a
        self.offset_b = offset_b

    def get_action(self, inventory):
        return np.array([self.offset_a, self.offset_b])

    def table(self, bundle, tick_size=0.001):
        """int32[T,5,2]: ``np.round(raw_offsets).astype(int)`` (main.py:113) for inv = -2..2."""
        T = len(bundle[2])
        act = np.round(self.get_action(0)).astype(np.int32)
        return np.broadcast_to(act, (T, 5, 2)).copy()


class GLFTPolicy:
    """Gueant-Lehalle-Fernandez-Tapia closed-form quotes around the mid (Env/benchmarks.py:16-40)."""

    def __init__(self, gamma=0.001, kappa=100, A=0.1, sigma=0.01):
        self.gamma = gamma
        self.kappa = kappa
        self.A = A
        self.sigma = sigma

    def get_action(self, inventory):
        g, k = self.gamma, self.kappa
        skew = np.sqrt((self.sigma ** 2 * g) / (2 * k * self.A) * (1 + g / k) ** (1 + k / g))
        half_spread = (1 / g) * np.log(1 + g / k)
        q = inventory
        ask_offset_mid = half_spread + ((2 * q - 1) / 2) * skew
        bid_offset_mid = half_spread - ((2 * q + 1) / 2) * skew
        return np.array([ask_offset_mid, bid_offset_mid])

    def table(self, bundle, tick_size=0.001):
        """int32[T,5,2]: the mid-relative offsets converted to best-relative ticks exactly as
        main.py:104-111 does (float64, same operation order), for inv = -2..2."""
        ask = np.asarray(bundle[3], np.float64)
        bid = np.asarray(bundle[4], np.float64)
        mid_p = (ask + bid) / 2.0
        out = np.empty((len(ask), 5, 2), np.int32)
        for iv in range(5):
            raw = self.get_action(iv - 2)
            off_a = ((mid_p + raw[0]) - ask) / tick_size
            off_b = (bid - (mid_p - raw[1])) / tick_size
            out[:, iv, 0] = np.round(off_a).astype(int)
            out[:, iv, 1] = np.round(off_b).astype(int)
        return out
