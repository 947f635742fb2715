import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """GPU tests fail loudly rather than skip when selected with -m gpu on a box without CUDA;
    under the default CPU run (-m "not gpu") they are deselected by the marker expression."""
    return


@pytest.fixture(scope="session")
def golden():
    class G:
        backtest = np.load(os.path.join(GOLDEN, "backtest_510300.npz"))
        ckpt = np.load(os.path.join(GOLDEN, "checkpoints.npz"))
        ref = np.load(os.path.join(GOLDEN, "ref_rollouts.npz"))
        tanh = np.load(os.path.join(GOLDEN, "tanh_threshold.npz"))
    return G


REF_CASES = ("plain", "fee", "arl", "arl_fee", "idle", "fresh")


def ref_case(ref, name):
    """Unpack one imported-reference case of tests/golden/ref_rollouts.npz."""
    bundle = tuple(ref[f"{name}.bundle.{k}"] for k in
                   ("s1", "s2", "mid_next", "best_ask", "best_bid", "buy_max", "sell_min"))
    stats = {k: ref[f"{name}.stats.{k}"][()] for k in ("s1_m", "s1_s", "s2_m", "s2_s")}
    d = dict(bundle=bundle, stats=stats, genomes=ref[f"{name}.genomes"],
             adv=ref[f"{name}.adv_genomes"] if bool(ref[f"{name}.use_arl"]) else None,
             fee=float(ref[f"{name}.fee"]), use_arl=bool(ref[f"{name}.use_arl"]),
             fitness=ref[f"{name}.fitness"], trades=ref[f"{name}.trades"],
             margin=ref[f"{name}.min_margin_ticks"],
             trace={k.split(".trace.")[1]: ref[k] for k in ref.files if k.startswith(f"{name}.trace.")},
             phi=float(ref["phi"]), tick=float(ref["tick"]))
    return d
