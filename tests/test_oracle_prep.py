"""CPU: the oracle's restatement of the two 'next' rows of SURVEY.md 8f against vectors made by the reference's
own code (oracle/make_golden_prep.py -> tests/golden/ref_prep.npz):
  * bundle-builder window reductions  pipeline/agent_trainer.py:47-73
  * strategy analytics                analytics/mm_analyzer.py:5-56
"""
import os

import numpy as np
import pytest

from oracle import oracle

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def prep():
    return np.load(os.path.join(HERE, "golden", "ref_prep.npz"))


def bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


def day_bundle(prep, i, windows):
    """One day of load_signals_bundle given a window-reduction implementation."""
    s1, s2 = prep[f"win.day{i}.s1_pred"], prep[f"win.day{i}.s2_pred"]
    n = min(len(s1), len(s2))                                                    # agent_trainer.py:47
    mid, ask, bid, bmax, smin = windows(prep[f"win.day{i}.askprice1"], prep[f"win.day{i}.bidprice1"],
                                        prep[f"win.day{i}.p_buy_max"], prep[f"win.day{i}.p_sell_min"],
                                        int(prep["win.event_step"]), n)
    return s1[-n:][:-1], s2[-n:][:-1], mid, ask, bid, bmax, smin                   # :48, :61-62


def assert_bundle_equals_reference(prep, windows):
    parts = [day_bundle(prep, i, windows) for i in range(int(prep["win.n_days"]))]
    names = ("s1", "s2", "mid", "ask", "bid", "buy_max", "sell_min")
    for k, name in enumerate(names):
        got = np.concatenate([p[k] for p in parts])
        want = prep[f"win.out.{name}"]
        assert got.shape == want.shape, name
        if name in ("s1", "s2"):
            assert np.array_equal(got, want), name
        else:
            assert np.array_equal(bits(got), bits(want)), name                  # bit-exact, NaN patterns included


def test_window_reductions_match_reference_bundle(prep):
    assert_bundle_equals_reference(prep, oracle.bundle_windows)
    assert np.isnan(prep["win.out.buy_max"]).any() and np.isnan(prep["win.out.sell_min"]).any()   # all-NaN windows are covered


def test_window_edge_cases():
    a = np.arange(100.0)
    z = oracle.bundle_windows(a, a, a, a, 19, 1)
    assert all(len(x) == 0 for x in z)
    z = oracle.bundle_windows(a, a, a, a, 19, 0)
    assert all(len(x) == 0 for x in z)
    with pytest.raises(ValueError):
        oracle.bundle_windows(a, a, a, a, 19, 7)               # only ceil(100/19) = 6 sampled events
    mid, ask, bid, bmax, smin = oracle.bundle_windows(a, a + 1, a, a, 19, 6)
    assert list(ask) == [0, 19, 38, 57, 76] and list(bmax) == [19, 38, 57, 76, 95] and list(smin) == [0, 19, 38, 57, 76]
    assert list(mid) == [19.5, 38.5, 57.5, 76.5, 95.5]


def test_analytics_match_reference_summary(prep):
    for name in prep["ana.cases"]:
        got = oracle.analytics(prep[f"ana.{name}.wealth"], prep[f"ana.{name}.inventory"], prep[f"ana.{name}.is_trade"])
        want = prep[f"ana.{name}.summary"]
        assert np.array_equal(bits(got), bits(want)), (name, got, want)
