"""GPU (-m gpu): the device versions of the two 'next' rows of SURVEY.md 8f through the C ABI, bit for bit against
the vectors made by the reference's own code (tests/golden/ref_prep.npz) and against the oracle on larger inputs."""
import os

import numpy as np
import pytest
import torch

from test_oracle_prep import assert_bundle_equals_reference, bits

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def sg():
    assert torch.cuda.is_available()
    import sgmm_b200
    return sgmm_b200


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    return oracle


@pytest.fixture(scope="module")
def prep():
    return np.load(os.path.join(HERE, "golden", "ref_prep.npz"))


def test_device_window_reductions_match_reference_bundle(sg, prep):
    assert_bundle_equals_reference(prep, lambda *a: sg.bundle_windows(*a))


def test_device_day_bundle_and_concat(sg, prep):
    days = []
    for i in range(int(prep["win.n_days"])):
        ev = {c: prep[f"win.day{i}.{c}"] for c in ("askprice1", "bidprice1", "p_buy_max", "p_sell_min")}
        days.append(sg.day_bundle(ev, prep[f"win.day{i}.s1_pred"], prep[f"win.day{i}.s2_pred"], int(prep["win.event_step"])))
    b = sg.concat_days(days)
    for k, name in enumerate(("s1", "s2", "mid", "ask", "bid", "buy_max", "sell_min")):
        assert np.array_equal(np.asarray(b[k]).view(np.uint8), prep[f"win.out.{name}"].view(np.uint8)), name


@pytest.mark.parametrize("E,step,n", [(200000, 19, 10526), (200000, 19, 2), (200000, 7, 1), (53, 19, 3), (5000, 1, 5000)])
def test_device_window_reductions_vs_oracle_large_and_edges(sg, orc, E, step, n):
    rng = np.random.default_rng(E + step + n)
    bid = np.round(3.4 + 0.001 * np.cumsum(rng.integers(-2, 3, E)), 3)
    ask = np.round(bid + 0.001 * rng.integers(1, 3, E), 3)
    bmax = np.where(rng.random(E) < 0.3, np.nan, np.round(ask + 0.001 * rng.integers(-2, 4, E), 3))
    smin = np.where(rng.random(E) < 0.3, np.nan, np.round(bid - 0.001 * rng.integers(-2, 4, E), 3))
    got = sg.bundle_windows(ask, bid, bmax, smin, step, n)
    want = orc.bundle_windows(ask, bid, bmax, smin, step, n)
    for g, w in zip(got, want):
        assert g.shape == w.shape and np.array_equal(bits(g), bits(w))


def test_device_window_argument_errors(sg):
    a = np.arange(100.0)
    with pytest.raises(sg.SgmmError):
        sg.bundle_windows(a, a, a, a, 19, 7)          # only 6 sampled events
    assert all(len(x) == 0 for x in sg.bundle_windows(a, a, a, a, 19, 0))


def test_device_analytics_match_reference_summary(sg, prep):
    import pandas as pd
    for name in prep["ana.cases"]:
        w, iv, tr = prep[f"ana.{name}.wealth"], prep[f"ana.{name}.inventory"], prep[f"ana.{name}.is_trade"]
        got = sg.population_summary(w, iv, tr)[0]
        assert np.array_equal(bits(got), bits(prep[f"ana.{name}.summary"])), (name, got, prep[f"ana.{name}.summary"])
        s = sg.StrategyAnalytics(pd.DataFrame({"wealth": w, "inventory": iv, "is_trade": tr})).summary_dict
        assert list(s) == ["Total PnL", "MAP (Risk)", "PnLMAP (Eff)", "Max DD", "Sharpe", "Trades"]
        assert s["Trades"] == int(prep[f"ana.{name}.summary"][5])


def test_device_analytics_batch_vs_oracle(sg, orc):
    rng = np.random.default_rng(3)
    B, T = 300, 1777
    w = np.cumsum(rng.standard_normal((B, T)) * 0.01, axis=1)
    iv = rng.integers(-2, 3, (B, T)).astype(np.int32)
    tr = rng.random((B, T)) < rng.random((B, 1))
    tr[0] = False; tr[1] = False; tr[1, 5] = True
    got = sg.population_summary(w, iv, tr)
    for b in range(B):
        want = orc.analytics(w[b], iv[b], tr[b])
        assert np.array_equal(bits(got[b]), bits(want)), (b, got[b], want)


def test_trace_kernel_outputs_feed_the_analytics(sg, orc):
    """Device trace -> recorder frame -> analytics: the golden ARL backtest end to end (notebook cell 34 figures
    are computed by the reference from these columns)."""
    b = np.load(os.path.join(HERE, "golden", "backtest_510300.npz"))
    for name in ("drl", "arl", "glft", "foic"):
        is_trade = (b[f"{name}.fill_buy"] != 0) | (b[f"{name}.fill_sell"] != 0)
        got = sg.population_summary(b[f"{name}.wealth"], b[f"{name}.inventory"], is_trade)[0]
        want = orc.analytics(b[f"{name}.wealth"], b[f"{name}.inventory"], is_trade)
        assert np.array_equal(bits(got), bits(want))
        assert got[5] == float(is_trade.sum())
