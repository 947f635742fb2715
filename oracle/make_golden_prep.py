"""Golden vectors for the two "next" rows of SURVEY.md 8f, produced by the REFERENCE'S OWN CODE run in the
build container (it cannot travel to the GPU box, so the vectors are committed):

  windows   pipeline/agent_trainer.py:15-77 load_signals_bundle, called unmodified.  Its collaborators that need
            packages or data this container lacks are replaced by fakes (xgboost / matplotlib module stubs, fake
            SGU models, fake loaders that hand back a prepared event_df, parquet files of one dummy row), so that
            the function's own sampling / window / gather code (:47-73) runs on synthetic event frames.
  analytics analytics/mm_analyzer.py:5-56 StrategyAnalytics.summary_dict on frames built from the shipped golden
            backtests (tests/golden/backtest_510300.npz) and on synthetic edge cases.

    python oracle/make_golden_prep.py        ->  tests/golden/ref_prep.npz
"""
import os
import sys
import tempfile
import types

import numpy as np
import pandas as pd

sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def stub_modules():
    for name in ("xgboost", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]


def synthetic_events(rng, E, nan_frac):
    bid = np.round(3.48 + 0.001 * np.cumsum(rng.integers(-2, 3, E)), 3)
    ask = np.round(bid + 0.001 * rng.integers(1, 3, E), 3)
    bmax = np.round(ask + 0.001 * rng.integers(-2, 4, E), 3)
    smin = np.round(bid - 0.001 * rng.integers(-2, 4, E), 3)
    bmax[rng.random(E) < nan_frac] = np.nan
    smin[rng.random(E) < nan_frac] = np.nan
    return pd.DataFrame({"askprice1": ask, "bidprice1": bid, "p_buy_max": bmax, "p_sell_min": smin})


def make_windows(out):
    stub_modules()
    sys.path.insert(0, REF)
    import pipeline.agent_trainer as at                      # the reference module, unmodified

    rng = np.random.default_rng(20240612)
    days = []
    # (events, signals from SGU1, signals from SGU2, NaN fraction): ragged tails, all-NaN windows, n == sampled count
    for (E, n1, n2, nf) in [(400, 18, 20, 0.05), (381, 21, 19, 0.6), (77, 5, 4, 1.0), (39, 3, 3, 0.0), (1000, 40, 53, 0.2)]:
        days.append((synthetic_events(rng, E, nf), rng.standard_normal(n1).astype(np.float32),
                     rng.standard_normal(n2).astype(np.float32)))
    state = {"day": -1}

    class FakeLoader1:
        def __init__(self, tick_df, snap_df):
            state["day"] += 1
        def gen_dataset(self, event_step):
            n1 = len(days[state["day"]][1])
            return pd.DataFrame({"f": np.zeros(n1), "label": np.zeros(n1)})

    class FakeLoader2:
        def __init__(self, tick_df, snap_df):
            self.event_df = days[state["day"]][0]
        def gen_dataset(self, event_step, time_steps):
            n2 = len(days[state["day"]][2])
            return np.zeros((n2, time_steps, 1), np.float32), np.zeros(n2, np.float32)

    class M1:
        def predict(self, X):
            return days[state["day"]][1]

    class M2:
        def predict(self, X):
            return days[state["day"]][2].reshape(-1, 1)

    class Scaler:
        def transform(self, X):
            return X

    at.SGU1DataPro, at.SGU2DataPro = FakeLoader1, FakeLoader2
    at.tqdm = lambda it, **kw: it
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "data", "SYN", "snap"))
        os.makedirs(os.path.join(tmp, "data", "SYN", "tick"))
        dates = [f"2024010{i}" for i in range(len(days))]
        dummy = pd.DataFrame({"trade_time": [93000001], "askprice1": [1.0], "bidprice1": [0.9]})
        for d in dates:
            dummy.to_parquet(os.path.join(tmp, "data", "SYN", "snap", f"{d}.parquet"))
            dummy.to_parquet(os.path.join(tmp, "data", "SYN", "tick", f"{d}.parquet"))
        os.chdir(tmp)
        try:
            bundle = at.load_signals_bundle("SYN", dates, M1(), M2(), Scaler())
        finally:
            os.chdir(cwd)
    names = ("s1", "s2", "mid", "ask", "bid", "buy_max", "sell_min")
    for k, v in zip(names, bundle):
        out[f"win.out.{k}"] = np.asarray(v)
    out["win.n_days"] = np.int64(len(days))
    for i, (df, s1, s2) in enumerate(days):
        for c in df.columns:
            out[f"win.day{i}.{c}"] = df[c].to_numpy()
        out[f"win.day{i}.s1_pred"] = s1
        out[f"win.day{i}.s2_pred"] = s2
    out["win.event_step"] = np.int64(19)


def make_analytics(out):
    stub_modules()
    sys.path.insert(0, REF)
    from analytics.mm_analyzer import StrategyAnalytics       # the reference class, unmodified
    b = np.load(os.path.join(ROOT, "tests", "golden", "backtest_510300.npz"))
    cases = {}
    for name in ("drl", "arl", "glft", "foic"):
        is_trade = (b[f"{name}.fill_buy"] != 0) | (b[f"{name}.fill_sell"] != 0)
        cases[name] = (b[f"{name}.wealth"], b[f"{name}.inventory"], is_trade)
    rng = np.random.default_rng(7)
    T = 300
    w = np.cumsum(rng.standard_normal(T) * 0.01)
    cases["no_trades"] = (w, rng.integers(-2, 3, T), np.zeros(T, bool))
    one = np.zeros(T, bool); one[17] = True
    cases["one_trade"] = (w, rng.integers(-2, 3, T), one)
    two = np.zeros(T, bool); two[[5, 200]] = True
    cases["two_trades"] = (w, np.zeros(T, np.int64), two)                      # MAP == 0 -> PnL/MAP = 0
    flat = np.zeros(T, bool); flat[::3] = True
    cases["flat_wealth"] = (np.full(T, 1.25), rng.integers(-2, 3, T), flat)   # std == 0 -> Sharpe = 0
    big = np.cumsum(rng.standard_normal(5000) * 0.003)
    cases["long"] = (big, rng.integers(-2, 3, 5000), rng.random(5000) < 0.8)
    keys = ("Total PnL", "MAP (Risk)", "PnLMAP (Eff)", "Max DD", "Sharpe", "Trades")
    out["ana.cases"] = np.array(sorted(cases))
    for name, (wealth, inv, tr) in cases.items():
        df = pd.DataFrame({"wealth": np.asarray(wealth, np.float64), "inventory": np.asarray(inv, np.int64), "is_trade": tr})
        s = StrategyAnalytics(df).summary_dict
        out[f"ana.{name}.wealth"] = df["wealth"].to_numpy()
        out[f"ana.{name}.inventory"] = df["inventory"].to_numpy().astype(np.int32)
        out[f"ana.{name}.is_trade"] = df["is_trade"].to_numpy()
        out[f"ana.{name}.summary"] = np.array([float(s[k]) for k in keys], np.float64)


if __name__ == "__main__":
    out = {}
    make_windows(out)
    make_analytics(out)
    path = os.path.join(ROOT, "tests", "golden", "ref_prep.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: (v.shape if hasattr(v, "shape") else v) for k, v in list(out.items())[:8]})
