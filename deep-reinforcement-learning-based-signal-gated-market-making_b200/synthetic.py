"""Synthetic "510300-shaped" bundles.

A bundle is the reference's 7-tuple of equal-length 1-D arrays
``(s1, s2, mid_next, best_ask, best_bid, buy_max, sell_min)``
(/root/reference/pipeline/agent_trainer.py:75-77).  The statistics below were read off the
reference's shipped 960-step test bundle (output/510300/*/backtest_0.0001.parquet): 240 bars per
day, tick 0.001, bid random walk on the tick grid, spread 1 tick (97.4 %) or 2, excursions of the
next-bar extreme trade prices beyond the touch, ~1 % NaN bounds (no trade on that side), s1 an
AR(1) float32 series, s2 iid float32.  Days are concatenated with no reset marker, exactly like
the reference's multi-day bundle.
"""
from __future__ import annotations

import numpy as np

BARS_PER_DAY = 240
TICK = 0.001

_BID_STEP_VALUES = np.array([0, 1, -1, 2, -2, 3, -3, 4, -4, 5, -5, 6, -6])
_BID_STEP_PROBS = np.array([0.31, 0.22, 0.22, 0.10, 0.10, 0.02, 0.02] + [0.01 / 6] * 6)
_BID_STEP_PROBS = _BID_STEP_PROBS / _BID_STEP_PROBS.sum()
# excursion e of the extreme trade price beyond the touch, in ticks:
#   P(e>=0)=.96  P(e>=1)=.45  P(e>=2)=.15  P(e>=3)=.04  P(e>=4)=0
_EXC_VALUES = np.array([-1, 0, 1, 2, 3])
_EXC_PROBS = np.array([0.04, 0.51, 0.30, 0.11, 0.04])


def synthetic_day(day: int, bars: int = BARS_PER_DAY, start_bid_ticks: int = 3480):
    """One trading day; returns the 7-tuple plus the closing bid (in ticks)."""
    rng = np.random.default_rng(20240529 + int(day))
    n = bars + 1                                    # one extra bar: mid_next needs bar t+1
    steps = rng.choice(_BID_STEP_VALUES, size=n, p=_BID_STEP_PROBS)
    bid_ticks = np.empty(n, dtype=np.int64)
    cur = int(start_bid_ticks)
    for i in range(n):                              # reflect into a sane band around 3.48
        cur += int(steps[i])
        if cur < 3300:
            cur = 3300 + (3300 - cur)
        if cur > 3700:
            cur = 3700 - (cur - 3700)
        bid_ticks[i] = cur
    spread = np.where(rng.random(n) < 0.974, 1, 2)
    ask_ticks = bid_ticks + spread
    bid = np.round(bid_ticks * TICK, 3)             # 3-dp float64, like price/10000 in the loaders
    ask = np.round(ask_ticks * TICK, 3)
    mid = (ask + bid) / 2.0
    ea = rng.choice(_EXC_VALUES, size=n, p=_EXC_PROBS)
    eb = rng.choice(_EXC_VALUES, size=n, p=_EXC_PROBS)
    buy_max = np.round((ask_ticks + ea) * TICK, 3)
    sell_min = np.round((bid_ticks - eb) * TICK, 3)
    buy_max[rng.random(n) < 0.01] = np.nan
    sell_min[rng.random(n) < 0.01] = np.nan
    # signals: s1 ~ AR(1) rho=.88 mean 2.24 sd .37 clipped >= 1.68 ; s2 ~ N(.04, .55^2)
    rho, mu, sd = 0.88, 2.24, 0.37
    eps = rng.standard_normal(n) * sd * np.sqrt(1.0 - rho * rho)
    s1 = np.empty(n)
    x = rng.standard_normal() * sd
    for i in range(n):
        x = rho * x + eps[i]
        s1[i] = mu + x
    s1 = np.maximum(s1, 1.68).astype(np.float32)
    s2 = (0.04 + 0.55 * rng.standard_normal(n)).astype(np.float32)
    out = (s1[:-1], s2[:-1], mid[1:].copy(), ask[:-1].copy(), bid[:-1].copy(),
           buy_max[:-1].copy(), sell_min[:-1].copy())
    return out, int(bid_ticks[-1])


def synthetic_bundle(n_days: int, first_day: int = 0, bars_per_day: int = BARS_PER_DAY):
    """``n_days`` concatenated days -> the reference's 7-tuple (T = n_days*bars_per_day)."""
    parts = [[] for _ in range(7)]
    start = 3480
    for d in range(first_day, first_day + n_days):
        day, start = synthetic_day(d, bars_per_day, start)
        for k in range(7):
            parts[k].append(day[k])
    return tuple(np.concatenate(p) for p in parts)


def train_stats_of(bundle):
    """The reference's expression, verbatim dtypes (pipeline/agent_trainer.py:125-129)."""
    s1_t, s2_t = bundle[0], bundle[1]
    return {
        's1_m': np.mean(s1_t), 's1_s': np.std(s1_t) + 1e-9,
        's2_m': np.mean(s2_t), 's2_s': np.std(s2_t) + 1e-9,
    }


def policy_like_genomes(count: int, hidden: int = 32, seed: int = 0, sigma: float = 0.05,
                        out_scale: float = 1.0, out_bias=(0.0, 0.0)):
    """Genomes shaped like the reference's population: an orthogonal-initialised master
    (models/model.py:18-21: gain 0.9, bias 0.05) plus sigma*N(0,1) children (model.py:65-71).
    ``out_scale`` / ``out_bias`` rescale the last layer so offsets span several ticks like the
    trained agents do (fresh masters quote ~0 everywhere).  numpy only (no torch dependency)."""
    rng = np.random.default_rng(seed)

    def orth(rows, cols):
        a = rng.standard_normal((max(rows, cols), min(rows, cols)))
        q, r = np.linalg.qr(a)
        q = q * np.sign(np.diag(r))
        q = q if rows >= cols else q.T
        return (0.9 * q[:rows, :cols]).astype(np.float32)

    H = hidden
    W1, W2, W3 = orth(H, 3), orth(H, H), orth(2, H)
    b1 = np.full(H, 0.05, np.float32)
    b2 = np.full(H, 0.05, np.float32)
    b3 = np.full(2, 0.05, np.float32)
    W3 = (W3 * np.float32(out_scale)).astype(np.float32)
    b3 = (b3 + np.asarray(out_bias, np.float32)).astype(np.float32)
    master = np.concatenate([W1.ravel(), b1, W2.ravel(), b2, W3.ravel(), b3]).astype(np.float32)
    noise = rng.standard_normal((count, master.size)).astype(np.float32) * np.float32(sigma)
    return master, (master[None, :] + noise).astype(np.float32)
