"""CPU: pin the oracle (oracle/sgmm_oracle.c) against every known answer the reference holds.

 1. the four shipped backtests  -> env arithmetic given (action, fills), bit-exact     
 2. the ARL checkpoint + notebook train_stats -> policy forward + x5 + round, 960/960   
 3. outputs of the imported reference on seeded synthetic bundles (oracle/make_golden.py)
 4. adversary tanh rounding threshold; degenerate episode lengths; argmax tie-break
"""
import numpy as np
import pytest

from conftest import REF_CASES, ref_case
from oracle import oracle


@pytest.mark.parametrize("name", ["drl", "arl", "glft", "foic"])
def test_backtest_env_arithmetic_bit_exact(golden, name):
    b = golden.backtest
    T = 960
    off = np.stack([b[f"{name}.off_a"], b[f"{name}.off_b"]], 1).astype(np.int32)
    fb, fs = b[f"{name}.fill_buy"], b[f"{name}.fill_sell"]
    # bounds are not recorded in the parquets: force the recorded fill decisions
    buy_max = np.where(fs == 1, np.inf, -np.inf)
    sell_min = np.where(fb == 1, -np.inf, np.inf)
    z = np.zeros(T, np.float32)
    bz = (z, z, b[f"{name}.mid"], b[f"{name}.ask"], b[f"{name}.bid"], buy_max, sell_min)
    fit, trades, tr = oracle.rollout(None, None, bz, 1e-4, 0.001, 0.0, forced_actions=off, trace=True)
    assert np.array_equal(tr["fill_buy"], fb) and np.array_equal(tr["fill_sell"], fs)
    assert np.array_equal(tr["inventory"], b[f"{name}.inventory"])
    for col in ("cash", "reward", "pnl_reward", "inventory_reward", "fee_paid"):
        assert np.array_equal(tr[col].view(np.uint64), b[f"{name}.{col}"].view(np.uint64)), col
    # recorder-derived columns (Env/recorder.py:45-51)
    assert np.array_equal(np.cumsum(tr["reward"]), b[f"{name}.cum_reward"])
    assert np.array_equal(tr["cash"] + tr["inventory"] * b[f"{name}.mid"], b[f"{name}.wealth"])
    assert trades == int(((fb == 1) | (fs == 1)).sum())


def test_backtest_aggregate_fills(golden):
    # notebook cell 32: fills ARL 1568, DRL 1612, GLFT 1825, FOIC 1821
    b = golden.backtest
    want = {"arl": 1568, "drl": 1612, "glft": 1825, "foic": 1821}
    for k, v in want.items():
        assert int(b[f"{k}.fill_buy"].sum() + b[f"{k}.fill_sell"].sum()) == v


def _arl_states(golden):
    b, c = golden.backtest, golden.ckpt
    s1, s2 = b["arl.s1_pred"], b["arl.s2_pred"]
    assert s1.dtype == np.float32
    z1 = ((s1 - c["train_stats_s1_m"][()]) / c["train_stats_s1_s"][()]).astype(np.float32)
    z2 = ((s2 - c["train_stats_s2_m"][()]) / c["train_stats_s2_s"][()]).astype(np.float32)
    inv_prev = np.concatenate([[0], b["arl.inventory"][:-1]])
    return z1, z2, inv_prev


def test_arl_checkpoint_actions_960_of_960(golden):
    b, c = golden.backtest, golden.ckpt
    g = c["510300_with_adv"]
    assert g.shape == (1250,) and g.dtype == np.float32
    z1, z2, inv_prev = _arl_states(golden)
    bad = 0
    min_margin = 1.0
    for t in range(960):
        raw = oracle.mlp_forward(g, [z1[t], z2[t], inv_prev[t] / 2.0])
        act = oracle.quantise(raw)
        q = raw * np.float32(5.0)
        min_margin = min(min_margin, float(np.min(np.abs(np.abs(q - np.floor(q)) - 0.5))))
        bad += int(act[0] != b["arl.off_a"][t]) + int(act[1] != b["arl.off_b"][t])
    assert bad == 0
    assert min_margin > 1e-4          # survey: 2.63e-4 ticks on this set


def test_arl_checkpoint_closed_loop_needs_bounds_free_replay(golden):
    """Closed loop on the golden bundle with fills forced from the parquet reproduces its rewards."""
    b, c = golden.backtest, golden.ckpt
    z1, z2, _ = _arl_states(golden)
    fb, fs = b["arl.fill_buy"], b["arl.fill_sell"]
    # bounds that yield the recorded fills for the recorded quotes (any consistent bounds do)
    my_ask = b["arl.ask"] + b["arl.off_a"] * 0.001
    my_bid = b["arl.bid"] - b["arl.off_b"] * 0.001
    buy_max = np.where(fs == 1, my_ask, my_ask - 0.0005)
    sell_min = np.where(fb == 1, my_bid, my_bid + 0.0005)
    # a blocked side (inventory cap) may have had a touch; keep it non-filling either way
    bz = (z1, z2, b["arl.mid"], b["arl.ask"], b["arl.bid"], buy_max, sell_min)
    fit, trades, tr = oracle.rollout(c["510300_with_adv"], None, bz, 1e-4, 0.001, 0.0, trace=True)
    assert np.array_equal(tr["off_a"], b["arl.off_a"]) and np.array_equal(tr["off_b"], b["arl.off_b"])
    assert np.array_equal(tr["inventory"], b["arl.inventory"])
    assert np.array_equal(tr["reward"], b["arl.reward"])
    assert fit == float(np.cumsum(b["arl.reward"])[-1]) or abs(fit - b["arl.cum_reward"][-1]) < 1e-12


def test_checkpoint_shape_contract(golden):
    for k in golden.ckpt.files:
        if not k.startswith("train_stats"):
            assert golden.ckpt[k].shape == (oracle.genome_len(32),) == (1250,)


@pytest.mark.parametrize("name", REF_CASES)
def test_oracle_vs_imported_reference(golden, name):
    cs = ref_case(golden.ref, name)
    z1, z2 = oracle.normalise(cs["bundle"], cs["stats"])
    bz = (z1, z2) + cs["bundle"][2:]
    P = cs["genomes"].shape[0]
    n_identical = 0
    for i in range(P):
        adv = cs["adv"][i] if cs["use_arl"] else None
        fit, trades, tr = oracle.rollout(cs["genomes"][i], adv, bz, cs["phi"], cs["tick"], cs["fee"],
                                         trace=True)
        ref_tr = {k: v[i] for k, v in cs["trace"].items()}
        # normalised inputs are bit-identical to what the reference fed torch
        assert np.array_equal(ref_tr["z1"], z1) and np.array_equal(ref_tr["z2"], z2)
        same_actions = (np.array_equal(tr["off_a"], ref_tr["off_a"]) and
                        np.array_equal(tr["off_b"], ref_tr["off_b"]) and
                        np.array_equal(tr["adv_a"], ref_tr["adv_a"]) and
                        np.array_equal(tr["adv_b"], ref_tr["adv_b"]))
        if same_actions:
            n_identical += 1
            # identical actions => integer work and every fp64 quantity are bit-exact
            for k in ("fill_buy", "fill_sell", "inventory"):
                assert np.array_equal(tr[k], ref_tr[k]), k
            for k in ("cash", "reward", "pnl_reward", "inventory_reward", "fee_paid"):
                assert np.array_equal(tr[k].view(np.uint64), ref_tr[k].view(np.uint64)), k
            assert trades == cs["trades"][i]
            assert fit == cs["fitness"][i]
            assert np.max(np.abs(tr["raw_a"] - ref_tr["raw_a"])) < 2e-5
        else:
            # a flipped rounding is only legitimate at a near-tie of the reference's own raw*5
            t = int(np.argmax((tr["off_a"] != ref_tr["off_a"]) | (tr["off_b"] != ref_tr["off_b"]) |
                              (tr["adv_a"] != ref_tr["adv_a"]) | (tr["adv_b"] != ref_tr["adv_b"])))
            q = np.array([ref_tr["raw_a"][t], ref_tr["raw_b"][t]], np.float32) * np.float32(5.0)
            assert np.min(np.abs(np.abs(q - np.floor(q)) - 0.5)) < 1e-4, (name, i, t, q)
    # SGMM-F32 order differs from MKL's by ~1e-7: essentially every trajectory is identical
    assert n_identical >= P - 1, (name, n_identical, P)


def test_degenerate_lengths(golden):
    import importlib.util, os
    ref = golden.ref
    assert ref["degenerate.T0"].tolist() == [-50.0, 0.0]
    z = np.zeros(0, np.float32)
    d = np.zeros(0)
    fit, trades = oracle.rollout(np.zeros(1250, np.float32), None, (z, z, d, d, d, d, d), 1e-4, 0.001, 0.0)
    assert (fit, trades) == (-50.0, 0)


def test_tanh_threshold_matches_torch(golden):
    thr = golden.tanh["thr"][()]
    assert thr == np.float32(0.54930615) and int(golden.tanh["bits"]) == 0x3F0C9F54
    g = np.zeros(1250, np.float32)
    g[36] = 1.0           # c1[0] = 1 -> h[0] = 1 for x = 0
    nxt = np.nextafter(thr, np.float32(1))
    for w, want in ((thr, 0), (nxt, 1), (-thr, 0), (-nxt, -1)):
        g[48] = w         # V2[0,0]
        pre, d = oracle.adv_forward(g, [0.0, 0.0, 0.0])
        assert pre[0] == w and d[0] == want


def test_argmax_first_max():
    assert oracle.argmax([1.0, 3.0, 3.0, 2.0]) == 1 == int(np.argmax([1.0, 3.0, 3.0, 2.0]))
    assert oracle.argmax([-50.0, -50.0]) == 0
    assert oracle.argmax([1.0, np.nan, 5.0]) == int(np.argmax([1.0, np.nan, 5.0]))


def test_mutation_is_standard_normal_and_counter_based():
    m = np.zeros(1250, np.float32)
    a = oracle.mutate(m, 1.0, seed=7, generation=3, individual=11)
    b = oracle.mutate(m, 1.0, seed=7, generation=3, individual=11)
    c = oracle.mutate(m, 1.0, seed=7, generation=3, individual=12)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    x = np.concatenate([oracle.mutate(m, 1.0, 1, 0, i) for i in range(400)])
    assert abs(x.mean()) < 0.01 and abs(x.std() - 1.0) < 0.01
    k = ((x - x.mean()) ** 4).mean() / x.var() ** 2
    assert abs(k - 3.0) < 0.05
    from scipy import stats
    assert stats.kstest(x[:20000], "norm").pvalue > 1e-3
    # child = master + sigma*noise, one fp32 mul + one fp32 add (models/model.py:69-70)
    mm = np.linspace(-1, 1, 1250).astype(np.float32)
    ch = oracle.mutate(mm, 0.05, 7, 3, 11)
    assert np.array_equal(ch, mm + a * np.float32(0.05))


# ----------------------------------------------------------------------------------------------
# long episodes: the unmodified reference at BASELINE config-2 length (oracle/make_golden_long.py)
# ----------------------------------------------------------------------------------------------
def _long_case(name):
    import os
    from conftest import GOLDEN
    from sgmm_b200 import synthetic
    g = np.load(os.path.join(GOLDEN, "ref_long.npz"))
    bundle = synthetic.synthetic_bundle(int(g["days"]))
    stats = synthetic.train_stats_of(bundle)
    _, genomes = synthetic.policy_like_genomes(int(g["P"]), seed=int(g["genome_seed"]), out_scale=float(g["out_scale"]),
                                               out_bias=tuple(g["out_bias"]))
    adv = (np.random.default_rng(int(g["adv_seed"])).standard_normal((int(g["P"]), 1250)) * float(g["adv_scale"])).astype(np.float32)
    use_arl = bool(g[f"{name}.use_arl"])
    return bundle, stats, genomes, (adv if use_arl else None), float(g[f"{name}.fee"]), g[f"{name}.fitness"], g[f"{name}.trades"]


def check_against_long_reference(name, fit, trd, oracle_mod, rel_tol=1e-5, near_tie_ticks=1e-4):
    """Trades identical and fitness within rel_tol of the reference's own output; a differing trajectory is excused only
    where the oracle's own raw*5 comes within near_tie_ticks of a rounding boundary.  Logs every trajectory's minimum margin
    (SURVEY.md 7.4-1)."""
    bundle, stats, genomes, adv, fee, f_ref, t_ref = _long_case(name)
    z1, z2 = oracle_mod.normalise(bundle, stats)
    bz = (z1, z2) + bundle[2:]
    excused = 0
    for i in range(len(f_ref)):
        _, _, tr = oracle_mod.rollout(genomes[i], None if adv is None else adv[i], bz, 1e-4, 0.001, fee, trace=True)
        q = np.stack([tr["raw_a"], tr["raw_b"]], 1).astype(np.float32) * np.float32(5.0)
        margin = float(np.min(np.abs(np.abs(q - np.floor(q)) - 0.5)))
        same = trd[i] == t_ref[i] and abs(fit[i] - f_ref[i]) <= rel_tol * max(1.0, abs(f_ref[i]))
        print(f"{name}[{i}]: min margin {margin:.3e} tick over {len(z1)} bars, trades {trd[i]} (reference {t_ref[i]}), "
              f"fitness {fit[i]:.9f} (reference {f_ref[i]:.9f}) -> {'identical' if same else 'differs'}")
        if not same:
            assert margin <= near_tie_ticks, (name, i, margin)
            excused += 1
    assert excused <= 2
    return excused


@pytest.mark.parametrize("name", ["plain", "arl", "fee", "arl_fee"])
def test_oracle_matches_the_reference_on_14400_bar_episodes(name):
    from oracle import oracle
    bundle, stats, genomes, adv, fee, f_ref, t_ref = _long_case(name)
    z1, z2 = oracle.normalise(bundle, stats)
    fit, trd = oracle.rollout_population((z1, z2) + bundle[2:], 1e-4, 0.001, fee, genomes=genomes, adv_genomes=adv, use_adv=adv is not None)
    check_against_long_reference(name, fit, trd, oracle)
