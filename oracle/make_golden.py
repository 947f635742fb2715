"""Generate tests/golden/*.npz by running the UNMODIFIED reference in this container.

Run from the repo root:   python oracle/make_golden.py
Needs /root/reference (read-only).  The GPU box does not have it, so the outputs are committed.

Fixtures written:
  backtest_510300.npz   the reference's four shipped backtests (output/510300/*/backtest_0.0001.parquet),
                        columns needed to replay Env/market_env.py:30-58
  checkpoints.npz       flat genomes of every shipped agent checkpoint + notebook cell-19 train_stats
  ref_rollouts.npz      seeded synthetic bundles + genomes + outputs of the imported reference:
                        evaluate_individual (Env/drl_engine.py:9-67) and a per-step trace made with
                        the reference's own FTPEnv / TradingPolicy / AdversaryPolicy objects
  tanh_threshold.npz    largest fp32 y with np.round(torch.tanh(y)) == 0 (adversary rounding)
"""
import importlib.util
import os
import sys

sys.dont_write_bytecode = True
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, REF)
from Env.market_env import FTPEnv                                   # noqa: E402
from Env.drl_engine import evaluate_individual                      # noqa: E402
from models.model import TradingPolicy, AdversaryPolicy             # noqa: E402

_spec = importlib.util.spec_from_file_location(
    "sgmm_synth", os.path.join(ROOT, "deep-reinforcement-learning-based-signal-gated-market-making_b200",
                               "synthetic.py"))
synth = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(synth)

torch.set_num_threads(1)
os.makedirs(OUT, exist_ok=True)


def state_dict_to_genome(sd, keys):
    return np.concatenate([sd[k].numpy().astype(np.float32).ravel() for k in keys])


def backtests():
    import pandas as pd
    cols = ["mid", "ask", "bid", "off_a", "off_b", "reward", "inventory", "cash", "fee_paid",
            "s1_pred", "s2_pred", "pnl_reward", "inventory_reward", "fill_buy", "fill_sell",
            "spread", "wealth", "cum_reward", "skew", "cum_fees", "realized_pnl", "unrealized_pnl"]
    out = {}
    for name in ("drl", "arl", "glft", "foic"):
        df = pd.read_parquet(f"{REF}/output/510300/{name}/backtest_0.0001.parquet")
        for c in cols:
            out[f"{name}.{c}"] = df[c].to_numpy()
    np.savez_compressed(os.path.join(OUT, "backtest_510300.npz"), **out)
    print("backtest_510300.npz", len(out), "arrays")


def checkpoints():
    keys = [f"net.{i}.{p}" for i in (0, 2, 4) for p in ("weight", "bias")]
    out = {}
    paths = {
        "510300_with_adv": "checkpoints/510300/with_adv/agent_best_val_0.0001.pth",
        "510300_without_adv": "checkpoints/510300/without_adv/agent_best_val_0.0001.pth",
        "688981_0.001": "checkpoints/688981/agent_best_val_0.001.pth",
        "688981_0.005": "checkpoints/688981/agent_best_val_0.005.pth",
        "688981_0.008": "checkpoints/688981/agent_best_val_0.008.pth",
        "688981_0.01": "checkpoints/688981/agent_best_val_0.01.pth",
    }
    for name, p in paths.items():
        sd = torch.load(os.path.join(REF, p), weights_only=True)
        assert list(sd.keys()) == keys, sd.keys()
        out[name] = state_dict_to_genome(sd, keys)
        # the flat layout must equal the reference's get_weights() (models/model.py:28-29)
        pol = TradingPolicy()
        pol.load_state_dict(sd)
        assert np.array_equal(pol.get_weights().numpy(), out[name])
    # notebook MM_replication_Report_JiaxingWei.ipynb cell 19 output (numpy-1.x dtypes: f32 mean, f64 std)
    out["train_stats_s1_m"] = np.float32(2.0911791)
    out["train_stats_s1_s"] = np.float64(0.3246540139184723)
    out["train_stats_s2_m"] = np.float32(0.027029233)
    out["train_stats_s2_s"] = np.float64(0.5160724530683288)
    np.savez_compressed(os.path.join(OUT, "checkpoints.npz"), **out)
    print("checkpoints.npz", list(out))


def traced_reference_loop(mm_w, adv_w, bundle, phi, tick, fee, stats, use_arl):
    """The reference's inner loop (Env/drl_engine.py:24-67) re-typed with its own objects, the way
    pipeline/evaluator.py and main.py:49-97 re-type it, recording every step."""
    s1, s2, mid_next, best_ask, best_bid, buy_max, sell_min = bundle
    policy = TradingPolicy()
    policy.set_weights(mm_w)
    adv_policy = None
    if use_arl and adv_w is not None:
        adv_policy = AdversaryPolicy()
        adv_policy.set_weights(adv_w)
    env = FTPEnv(phi=phi, tick_size=tick, fee_rate=fee)
    T = len(mid_next)
    rec = {k: np.zeros(T, np.int32) for k in ("off_a", "off_b", "adv_a", "adv_b", "fill_buy",
                                               "fill_sell", "inventory")}
    rec.update({k: np.zeros(T, np.float64) for k in ("cash", "reward", "pnl_reward",
                                                      "inventory_reward", "fee_paid")})
    rec.update({k: np.zeros(T, np.float32) for k in ("raw_a", "raw_b", "z1", "z2")})
    total_reward, trades = 0, 0
    fb, fs = 0.0, 0.0
    with torch.no_grad():
        for t in range(T):
            st = torch.tensor([[(s1[t] - stats['s1_m']) / stats['s1_s'],
                                (s2[t] - stats['s2_m']) / stats['s2_s'],
                                env.inventory / 2.0]], dtype=torch.float32)
            raw = policy.forward(st).squeeze().cpu().numpy()
            act = np.round(raw * 5.0).astype(int)
            adv_action = None
            if adv_policy is not None:
                ast = torch.tensor([[env.inventory / 2.0, fs, fb]], dtype=torch.float32)
                adv_raw = adv_policy.forward(ast).squeeze().cpu().numpy()
                adv_action = np.round(adv_raw * 1.0).astype(int)
            reward, info = env.step(act, mid_next[t], best_ask[t], best_bid[t], buy_max[t],
                                    sell_min[t], adv_action=adv_action)
            total_reward += reward
            fb = 1.0 if info['fill_buy'] else 0.0
            fs = 1.0 if info['fill_sell'] else 0.0
            if info['fill_buy'] or info['fill_sell']:
                trades += 1
            rec["z1"][t], rec["z2"][t] = st[0, 0].item(), st[0, 1].item()
            rec["raw_a"][t], rec["raw_b"][t] = raw[0], raw[1]
            rec["off_a"][t], rec["off_b"][t] = act[0], act[1]
            if adv_action is not None:
                rec["adv_a"][t], rec["adv_b"][t] = adv_action[0], adv_action[1]
            rec["fill_buy"][t], rec["fill_sell"][t] = info['fill_buy'], info['fill_sell']
            rec["inventory"][t], rec["cash"][t] = env.inventory, env.cash
            rec["reward"][t], rec["pnl_reward"][t] = reward, info['pnl_reward']
            rec["inventory_reward"][t], rec["fee_paid"][t] = info['inventory_reward'], info['fee_paid']
    if trades == 0:
        total_reward -= 50.0
    return float(total_reward), int(trades), rec


def make_adv_genomes(count, seed, scale):
    """Adversary genomes: the first 74 floats of a 1250-float TradingPolicy-shaped genome
    (models/model.py:52-57,63).  Scaled so that tanh outputs cross +-0.5 for some of the 20 states
    (fresh masters give all-zero displacements)."""
    rng = np.random.default_rng(seed)
    g = rng.standard_normal((count, 1250)).astype(np.float32) * np.float32(scale)
    return g


def ref_rollouts():
    out = {}
    cases = [
        # name, days, first_day, P, fee, use_arl, out_scale, out_bias, seed
        ("plain", 2, 0, 12, 0.0, False, 6.0, (0.15, 0.15), 1),
        ("fee", 2, 2, 8, 3e-4, False, 6.0, (0.15, 0.15), 2),
        ("arl", 2, 4, 12, 0.0, True, 6.0, (0.15, 0.15), 3),
        ("arl_fee", 1, 6, 6, 3e-5, True, 5.0, (0.1, 0.2), 4),
        ("idle", 1, 7, 4, 0.0, False, 1.0, (3.0, 3.0), 5),     # quotes 15 ticks away: no fill -> -50
        ("fresh", 4, 8, 8, 0.0, False, 1.0, (0.0, 0.0), 6),     # reference-initialised scale
    ]
    phi, tick = 1e-4, 0.001
    for name, days, d0, P, fee, arl, osc, ob, seed in cases:
        bundle = synth.synthetic_bundle(days, first_day=d0)
        stats = synth.train_stats_of(bundle)
        _, genomes = synth.policy_like_genomes(P, 32, seed=seed, sigma=0.05, out_scale=osc, out_bias=ob)
        advg = make_adv_genomes(P, seed + 100, 1.0) if arl else None
        fit = np.zeros(P)
        trd = np.zeros(P, np.int32)
        margin = np.zeros(P)
        recs = []
        for i in range(P):
            mmw = torch.from_numpy(genomes[i].copy())
            advw = torch.from_numpy(advg[i].copy()) if arl else None
            f, n = evaluate_individual(mmw, advw, bundle, phi, tick, fee, stats, use_arl=arl)
            f2, n2, rec = traced_reference_loop(mmw, advw, bundle, phi, tick, fee, stats, arl)
            assert n == n2 and (f == f2 or (np.isnan(f) and np.isnan(f2))), (name, i, f, f2, n, n2)
            fit[i], trd[i] = f, n
            q = np.stack([rec["raw_a"], rec["raw_b"]]).astype(np.float32) * np.float32(5.0)
            margin[i] = np.min(np.abs(np.abs(q - np.floor(q)) - 0.5))
            recs.append(rec)
        for k, v in zip(("s1", "s2", "mid_next", "best_ask", "best_bid", "buy_max", "sell_min"), bundle):
            out[f"{name}.bundle.{k}"] = v
        for k in ("s1_m", "s1_s", "s2_m", "s2_s"):
            out[f"{name}.stats.{k}"] = np.asarray(stats[k])
        out[f"{name}.genomes"] = genomes
        if arl:
            out[f"{name}.adv_genomes"] = advg
        out[f"{name}.fee"] = np.float64(fee)
        out[f"{name}.use_arl"] = np.bool_(arl)
        out[f"{name}.fitness"] = fit
        out[f"{name}.trades"] = trd
        out[f"{name}.min_margin_ticks"] = margin
        for k in recs[0]:
            out[f"{name}.trace.{k}"] = np.stack([r[k] for r in recs])
        print(name, "T", len(bundle[0]), "P", P, "fitness", np.round(fit[:4], 4), "trades", trd[:4],
              "min margin", margin.min(), "adv nonzero",
              int((out[f"{name}.trace.adv_a"] != 0).sum() + (out[f"{name}.trace.adv_b"] != 0).sum()))
    out["phi"] = np.float64(phi)
    out["tick"] = np.float64(tick)
    # degenerate episode lengths (T=0 -> (-50.0, 0); T=1)
    b1 = tuple(a[:1] for a in synth.synthetic_bundle(1, first_day=9))
    st = synth.train_stats_of(synth.synthetic_bundle(1, first_day=9))
    g = synth.policy_like_genomes(1, 32, seed=7, out_scale=6.0)[1][0]
    f0, n0 = evaluate_individual(torch.from_numpy(g.copy()), None, tuple(a[:0] for a in b1), phi, tick, 0.0, st)
    f1, n1 = evaluate_individual(torch.from_numpy(g.copy()), None, b1, phi, tick, 0.0, st)
    out["degenerate.T0"] = np.array([f0, n0], np.float64)
    out["degenerate.T1"] = np.array([f1, n1], np.float64)
    np.savez_compressed(os.path.join(OUT, "ref_rollouts.npz"), **out)
    print("ref_rollouts.npz", len(out), "arrays;", "T0 ->", f0, n0, " T1 ->", f1, n1)


def benchmark_policies():
    """FOIC / GLFT through the reference's own FTPEnv and Env/benchmarks.py classes with the driver
    loop of main.py:99-132 re-typed (main.py itself needs seaborn/statsmodels and the un-shipped data)."""
    from Env.benchmarks import FOICPolicy, GLFTPolicy
    out = {}
    bundle = synth.synthetic_bundle(2, first_day=11)
    s1, s2, mid, ask, bid, b_max, s_min = bundle
    for k, v in zip(("s1", "s2", "mid_next", "best_ask", "best_bid", "buy_max", "sell_min"), bundle):
        out[f"bundle.{k}"] = v
    pols = {"glft": GLFTPolicy(gamma=0.0001, kappa=3000, A=0.1, sigma=0.0005),       # main.py:206
            "glft_wide": GLFTPolicy(gamma=0.01, kappa=1500, A=0.1, sigma=0.02),
            "foic": FOICPolicy(offset_a=0, offset_b=0),                               # main.py:209
            "foic_1_2": FOICPolicy(offset_a=1, offset_b=2)}
    for name, policy in pols.items():
        for fee in (0.0, 3e-4):
            env = FTPEnv(phi=1e-4, tick_size=0.001, fee_rate=fee)
            T = len(mid)
            rec = {k: np.zeros(T, np.int32) for k in ("off_a", "off_b", "fill_buy", "fill_sell", "inventory")}
            rec.update({k: np.zeros(T, np.float64) for k in ("cash", "reward", "pnl_reward", "fee_paid")})
            for t in range(T):
                raw_offsets = policy.get_action(env.inventory)
                mid_p = (ask[t] + bid[t]) / 2.0
                if name.startswith("glft"):
                    off_a = ((mid_p + raw_offsets[0]) - ask[t]) / 0.001
                    off_b = (bid[t] - (mid_p - raw_offsets[1])) / 0.001
                    action = np.round([off_a, off_b]).astype(int)
                else:
                    action = np.round(raw_offsets).astype(int)
                reward, info = env.step(action, mid[t], ask[t], bid[t], b_max[t], s_min[t])
                rec["off_a"][t], rec["off_b"][t] = action[0], action[1]
                rec["fill_buy"][t], rec["fill_sell"][t] = info["fill_buy"], info["fill_sell"]
                rec["inventory"][t], rec["cash"][t] = env.inventory, env.cash
                rec["reward"][t], rec["pnl_reward"][t], rec["fee_paid"][t] = reward, info["pnl_reward"], info["fee_paid"]
            for k, v in rec.items():
                out[f"{name}.fee{fee}.{k}"] = v
            print("benchmark", name, fee, "fills", int(rec["fill_buy"].sum() + rec["fill_sell"].sum()),
                  "offsets", np.unique(rec["off_a"]), np.unique(rec["off_b"]))
    np.savez_compressed(os.path.join(OUT, "ref_benchmarks.npz"), **out)


def tanh_threshold():
    cur = np.float32(0.5493)
    ys = []
    for _ in range(4000):
        ys.append(cur)
        cur = np.nextafter(cur, np.float32(1))
    ys = np.array(ys, np.float32)
    r = np.round(torch.tanh(torch.from_numpy(ys)).numpy())
    assert np.all(np.diff(r) >= 0) and r[0] == 0 and r[-1] == 1
    i = int(np.argmax(r == 1))
    thr = ys[i - 1]
    # the same through a 1-row AdversaryPolicy-shaped call (nn.Tanh on a [1,2] tensor)
    t2 = np.round(torch.nn.Tanh()(torch.tensor([[float(thr), float(ys[i])]], dtype=torch.float32)).numpy())
    assert t2.tolist() == [[0.0, 1.0]]
    rn = np.round(torch.tanh(torch.from_numpy(-ys)).numpy())
    assert int(np.argmax(rn == -1)) == i
    np.savez(os.path.join(OUT, "tanh_threshold.npz"), thr=thr, bits=thr.view(np.uint32))
    print("tanh threshold", repr(thr), hex(int(thr.view(np.uint32))))


if __name__ == "__main__":
    benchmark_policies()
    backtests()
    checkpoints()
    tanh_threshold()
    ref_rollouts()
