"""GPU (-m gpu): the reference's unchanged `pipeline/agent_trainer.py::run_agent_training_pipeline` running on the B200
backend through `dropin/` (module shadowing, zero edits) -- the drop-in claim of INTEGRATION.md, exercised.

The caller is the reference's own file, staged byte for byte under oracle/_ref (oracle/stage_ref.py; git-ignored, travels to
the GPU box).  SGU models, loaders, parquet data and the plot are faked (tests/dropin_pipeline_driver.py), as in
oracle/make_golden_prep.py.  The run covers: load_signals_bundle x3, train_stats, DRLEngine(pop 50, use_arl).train for the
reference's hard-coded 100 generations, checkpoint save + reload, the per-bar blind-test loop over FTPEnv.step /
TradingPolicy.forward / StrategyRecorder.record, to_dataframe, StrategyAnalytics; then pipeline/evaluator.py::run_drl_backtest
(record_detailed + parquet) against the same backtest done as one device launch.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_reference_agent_training_pipeline_runs_unchanged_on_the_b200_backend(tmp_path):
    assert torch.cuda.is_available()
    staged = os.path.join(ROOT, "oracle", "_ref", "pipeline", "agent_trainer.py")
    if not os.path.exists(staged):
        pytest.skip("oracle/_ref/pipeline/agent_trainer.py is not staged on this box (python oracle/stage_ref.py where /root/reference exists)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin_pipeline_driver.py"), ROOT], cwd=tmp_path,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    info = json.loads(r.stdout.strip().splitlines()[-1])
    # the caller is the reference's file; what it imported is this repository's backend
    assert info["agent_trainer"].endswith(os.path.join("oracle", "_ref", "pipeline", "agent_trainer.py"))
    for m, f in info["modules"].items():
        assert os.sep + "dropin" + os.sep in f, (m, f)
    assert info["engine_class"] == "sgmm_b200.engine.DRLEngine" and info["env_class"] == "sgmm_b200.env.FTPEnv"
    assert info["days_loaded"] == 24 and info["csv_exists"] and info["checkpoint_exists"]
    assert "Gen 000 | ARL:ON | Best Train:" in r.stdout and "Gen 095" in r.stdout      # drl_engine.py:169-171 log lines
    # the blind test's rows: the recorder contract, and the device trace of the saved agent reproduces them bit for bit
    import pandas as pd
    import sgmm_b200
    df = pd.read_csv(info["csv"])
    for c in ("step", "mid", "ask", "bid", "off_a", "off_b", "reward", "inventory", "cash", "pnl_reward", "inventory_reward",
              "fee_paid", "is_trade", "spread", "wealth", "cum_reward", "skew", "cum_fees", "realized_pnl", "unrealized_pnl"):
        assert c in df.columns, c
    n_test = 4 * 59                                   # 24 S3 days -> 16 train / 4 val / 4 test, 59 bars per day
    assert len(df) == n_test == info["plotted"]["rows"]
    assert df["inventory"].abs().max() <= 2
    assert info["plotted"]["metrics"]["Trades"] == int(df["is_trade"].sum())
    # pipeline/evaluator.py::run_drl_backtest (unmodified) ran its per-bar loop on FTPEnv / TradingPolicy / StrategyRecorder.
    # record_detailed and wrote its parquet; the same backtest as ONE device launch (rollout_trace + StrategyRecorder.from_trace)
    # gives the same frame column for column, bit for bit (offsets, fills, inventory, cash, rewards, fees, derived columns)
    assert info["evaluator"].endswith(os.path.join("oracle", "_ref", "pipeline", "evaluator.py"))
    assert info["backtest_rows"] == 480 and os.path.exists(info["backtest_parquet"])
    assert all(info["backtest_device_trace_matches"].values()), info["backtest_device_trace_matches"]
    assert info["backtest_fitness"] == info["backtest_cum_reward"]
    sd = torch.load(info["checkpoint"], weights_only=True)
    assert list(sd.keys()) == [f"net.{i}.{p}" for i in (0, 2, 4) for p in ("weight", "bias")]
