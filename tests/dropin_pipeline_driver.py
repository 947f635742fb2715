"""Driver of tests/test_gpu_dropin_pipeline.py -- run as a script in a scratch working directory.

Runs the reference's OWN `pipeline/agent_trainer.py::run_agent_training_pipeline` (staged, unmodified, under oracle/_ref)
with this repository's `dropin/` first on sys.path, so that its `from Env.market_env import FTPEnv`,
`from Env.drl_engine import DRLEngine`, `from Env.recorder import StrategyRecorder` and `from models.model import
TradingPolicy` bind to the B200 backend.  What the pipeline needs from outside the hot path is faked exactly the way
oracle/make_golden_prep.py fakes it for the golden vectors: SGU models, loaders and parquet files that hand back prepared
synthetic event frames, and a no-op plot.  Prints one JSON line describing what ran.
"""
import json
import os
import sys
import types

ROOT, SYMBOL, PHI = sys.argv[1], "SYN", 0.0001
DROPIN = os.path.join(ROOT, "deep-reinforcement-learning-based-signal-gated-market-making_b200", "dropin")
sys.path[:0] = [DROPIN, ROOT, os.path.join(ROOT, "oracle", "_ref")]      # dropin shadows Env.* / models.model of the reference
sys.dont_write_bytecode = True

import numpy as np          # noqa: E402
import pandas as pd         # noqa: E402

N_DAYS, EVENTS, STEP = 8, 19 * 60 + 1, 19
rng = np.random.default_rng(20241018)


def synthetic_events(E):
    bid = np.round(3.48 + 0.001 * np.cumsum(rng.integers(-2, 3, E)), 3)
    ask = np.round(bid + 0.001 * rng.integers(1, 3, E), 3)
    bmax = np.round(ask + 0.001 * rng.integers(-1, 4, E), 3)
    smin = np.round(bid - 0.001 * rng.integers(-1, 4, E), 3)
    bmax[rng.random(E) < 0.02] = np.nan
    smin[rng.random(E) < 0.02] = np.nan
    return pd.DataFrame({"askprice1": ask, "bidprice1": bid, "p_buy_max": bmax, "p_sell_min": smin})


days = [(synthetic_events(EVENTS), (2.2 + 0.3 * rng.standard_normal(60)).astype(np.float32),
         (0.04 + 0.5 * rng.standard_normal(60)).astype(np.float32)) for _ in range(3 * N_DAYS)]
state = {"day": -1}


class SGU1:
    def load(self, path): self.path = path
    def predict(self, X): return days[state["day"]][1]


class SGU2:
    def __init__(self, input_size=1, hidden_size=10): pass
    def load(self, path): self.path = path
    def predict(self, X): return days[state["day"]][2].reshape(-1, 1)


class SGU1DataPro:
    def __init__(self, tick_df, snap_df): state["day"] += 1
    def gen_dataset(self, event_step):
        n = len(days[state["day"]][1])
        return pd.DataFrame({"f": np.zeros(n), "label": np.zeros(n)})


class SGU2DataPro:
    def __init__(self, tick_df, snap_df): self.event_df = days[state["day"]][0]
    def gen_dataset(self, event_step, time_steps):
        n = len(days[state["day"]][2])
        return np.zeros((n, time_steps, 1), np.float32), np.zeros(n, np.float32)


class Scaler:
    def transform(self, X): return X


def fake_module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


plotted = {}


class BacktestVisualizer:
    @staticmethod
    def plot_professional_report(df, metrics, save_path=None, show_fees=False, nb_mode=False):
        plotted["rows"], plotted["metrics"] = len(df), {k: float(v) for k, v in metrics.items()}


import torch                # noqa: E402
torch.manual_seed(20241018)  # the engine draws its Philox key and the initial policies from torch's global generator: make the run repeatable
import sgmm_b200            # noqa: E402  (the analytics class of this repository mirrors analytics/mm_analyzer.py)
fake_module("models.GateUnits", SGU1=SGU1, SGU2=SGU2)
fake_module("loaders")
fake_module("loaders.HFTLoader", SGU1DataPro=SGU1DataPro, SGU2DataPro=SGU2DataPro)
fake_module("analytics")
fake_module("analytics.mm_analyzer", StrategyAnalytics=sgmm_b200.StrategyAnalytics, BacktestVisualizer=BacktestVisualizer)
if "tqdm" not in sys.modules:
    try:
        import tqdm  # noqa: F401
    except ImportError:
        fake_module("tqdm", tqdm=lambda it, **kw: it)

# data/ and checkpoints/ as the pipeline expects them (cwd is a scratch directory)
import pickle               # noqa: E402
os.makedirs(f"data/{SYMBOL}/snap"); os.makedirs(f"data/{SYMBOL}/tick"); os.makedirs(f"checkpoints/{SYMBOL}")
dummy = pd.DataFrame({"trade_time": [93000001], "askprice1": [1.0], "bidprice1": [0.9]})
dates = [f"202406{d:02d}" for d in range(1, 3 * N_DAYS + 1)]
for d in dates:
    dummy.to_parquet(f"data/{SYMBOL}/snap/{d}.parquet")
    dummy.to_parquet(f"data/{SYMBOL}/tick/{d}.parquet")
with open(f"checkpoints/{SYMBOL}/sgu2_scaler_20240401_20240528.pkl", "wb") as f:
    pickle.dump(None, f)

import pipeline.agent_trainer as at        # noqa: E402  the reference's module, unmodified (oracle/_ref)
at.pickle = types.SimpleNamespace(load=lambda f: Scaler())      # the scaler object is a host-side sklearn artefact
import Env.drl_engine, Env.market_env, Env.recorder, models.model      # noqa: E402,E401

at.run_agent_training_pipeline(SYMBOL, (20240401, 20240528), PHI=PHI, TICK_SIZE=0.001, USE_FEE=False, USE_ARL=True)

# ---- the other caller: pipeline/evaluator.py::run_drl_backtest (the per-bar backtest loop with record_detailed + parquet) on a
# synthetic two-day bundle with the agent just trained; then the SAME backtest as one device launch (rollout_trace)
import pipeline.evaluator as ev             # noqa: E402  the reference's module, unmodified (oracle/_ref)
from sgmm_b200 import synthetic             # noqa: E402
bt_bundle = synthetic.synthetic_bundle(2, first_day=900)
bt_stats = synthetic.train_stats_of(bt_bundle)
weight_path = f"checkpoints/{SYMBOL}/with_adv/agent_best_val_{PHI}.pth"
ev.run_drl_backtest(SYMBOL, "arl", weight_path, bt_bundle, PHI, 0.0003, bt_stats)
bt_parquet = f"output/{SYMBOL}/arl/backtest_{PHI}.parquet"
pol = sgmm_b200.TradingPolicy(); pol.load_state_dict(torch.load(weight_path, weights_only=True))
bun = sgmm_b200.Bundle.from_arrays(bt_bundle, bt_stats, 0.001)
fit, trades, tr = sgmm_b200.rollout_trace(bun, pol.get_weights().numpy(), phi=PHI, fee_rate=0.0003)
dev_df = sgmm_b200.StrategyRecorder.from_trace(tr, bt_bundle).to_dataframe()
ref_df = pd.read_parquet(bt_parquet)
same_cols = {}
for c in ("off_a", "off_b", "fill_buy", "fill_sell", "inventory", "cash", "reward", "pnl_reward", "inventory_reward", "fee_paid",
          "spread", "wealth", "cum_reward", "skew", "cum_fees", "realized_pnl", "unrealized_pnl"):
    same_cols[c] = bool(np.array_equal(np.asarray(dev_df[c]), np.asarray(ref_df[c])))

csv = f"output/{SYMBOL}/phi_{PHI}_S3_TEST_results.csv"
ck = f"checkpoints/{SYMBOL}/with_adv/agent_best_val_{PHI}.pth"
print(json.dumps({
    "agent_trainer": at.__file__,
    "modules": {m: sys.modules[m].__file__ for m in ("Env.drl_engine", "Env.market_env", "Env.recorder", "models.model")},
    "engine_class": f"{at.DRLEngine.__module__}.{at.DRLEngine.__name__}",
    "env_class": f"{at.FTPEnv.__module__}.{at.FTPEnv.__name__}",
    "csv": os.path.abspath(csv), "csv_exists": os.path.exists(csv), "checkpoint": os.path.abspath(ck),
    "checkpoint_exists": os.path.exists(ck), "plotted": plotted, "days_loaded": state["day"] + 1,
    "evaluator": ev.__file__, "backtest_parquet": os.path.abspath(bt_parquet), "backtest_rows": int(len(ref_df)),
    "backtest_device_trace_matches": same_cols, "backtest_fitness": float(fit), "backtest_cum_reward": float(ref_df["cum_reward"].iloc[-1]),
}))
