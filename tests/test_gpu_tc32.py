"""GPU (-m gpu): the H=32 tensor-core rollout (sgmm_tc32.cu: all three policy layers on tcgen05, A operands
chained through tensor memory, bf16 inputs / fp32 accumulate).  Parity is stated in two halves, as for
the H=256 kernel:

  (1) POLICY OUTPUTS vs the fp32 oracle (SGMM-F32 order, oracle/sgmm_oracle.c) for EVERY (bar,
      inventory) pair: |d(raw*5)| <= TAU_TICKS, and the rounded offsets are identical wherever the
      oracle's own distance to a rounding boundary exceeds TAU_TICKS;
  (2) GIVEN the offsets the kernel took, fills, inventory, trade count, rewards and fitness are
      BIT-IDENTICAL to the oracle's env (teacher-forced replay of the kernel's action trace).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

# stated tolerance on raw*5 (ticks) for offsets spanning +-10 ticks, per precision; the measured maximum is printed
TAU = {"bf16": 0.12, "tf32": 0.03, "f16": 0.03}
TAU_TICKS = TAU["bf16"]


@pytest.fixture(scope="module")
def sg():
    assert torch.cuda.is_available()
    import sgmm_b200
    return sgmm_b200


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    return oracle


def _setup(sg, orc, days, first_day, P, seed, T=None, out_scale=6.0):
    from sgmm_b200 import synthetic
    bundle = synthetic.synthetic_bundle(days, first_day=first_day)
    if T is not None:
        bundle = tuple(a[:T] for a in bundle)
    stats = synthetic.train_stats_of(synthetic.synthetic_bundle(days, first_day=first_day))
    z1, z2 = orc.normalise(bundle, stats)
    bun = sg.Bundle.from_arrays(bundle, stats, 0.001)
    master, genomes = synthetic.policy_like_genomes(P, hidden=32, seed=seed, out_scale=out_scale, out_bias=(0.1, 0.1))
    return bundle, (z1, z2) + bundle[2:], bun, master, genomes


def _audit(orc, bz, genomes, raw, T, TAU_TICKS=TAU_TICKS):
    worst, flips_outside = 0.0, 0
    for i in range(genomes.shape[0]):
        for t in range(T):
            for iv in range(5):
                want = orc.mlp_forward(genomes[i], [bz[0][t], bz[1][t], (iv - 2) / 2.0], hidden=32)
                q_o = want * np.float32(5.0)
                q_k = raw[i, t, iv] * np.float32(5.0)
                worst = max(worst, float(np.max(np.abs(q_o - q_k))))
                margin = np.abs(np.abs(q_o - np.floor(q_o)) - 0.5)
                flips_outside += int(np.sum((np.rint(q_o) != np.rint(q_k)) & (margin > TAU_TICKS)))
    return worst, flips_outside


@pytest.mark.parametrize("precision", ["bf16", "tf32", "f16"])
@pytest.mark.parametrize("fee", [0.0, 3e-4])
def test_policy_outputs_within_tolerance_and_env_bit_exact(sg, orc, fee, precision):
    TAU_TICKS = TAU[precision]
    bundle, bz, bun, master, genomes = _setup(sg, orc, 1, 90, 5, seed=11)
    T = bun.T
    fit, trd, raw, act = sg.rollout_tc_audit(bun, genomes, phi=1e-4, fee_rate=fee, hidden=32, precision=precision)
    fit, trd, raw, act = fit.cpu().numpy(), trd.cpu().numpy(), raw.cpu().numpy(), act.cpu().numpy()
    worst, flips_outside = _audit(orc, bz, genomes, raw, T, TAU_TICKS)
    for i in range(genomes.shape[0]):
        fo, to, tro = orc.rollout(None, None, bz, 1e-4, 0.001, fee, forced_actions=act[i], trace=True)
        assert fo == fit[i] and to == trd[i], (i, fo, fit[i], to, trd[i])
        inv_before = np.concatenate([[0], tro["inventory"][:-1]])
        taken = np.rint(raw[i, np.arange(T), inv_before + 2] * np.float32(5.0)).astype(np.int32)
        assert np.array_equal(taken, act[i])
    print(f"{precision}: max |d(raw*5)| = {worst:.4g} ticks (tolerance {TAU_TICKS})")
    assert worst <= TAU_TICKS
    assert flips_outside == 0


@pytest.mark.parametrize("T,P,group", [(1, 3, 0), (24, 40, 0), (25, 150, 0), (26, 17, 16), (51, 33, 6), (130, 5, 2), (240, 300, 0)])
def test_ragged_lengths_groups_and_many_individuals(sg, orc, T, P, group):
    """Tail chunks, partial groups (P not a multiple of the group), several groups per CTA."""
    precision = "tf32" if (T % 2) else "bf16"
    bundle, bz, bun, master, genomes = _setup(sg, orc, 1, 91, P, seed=T, T=T)
    fit, trd, raw, act = sg.rollout_tc_audit(bun, genomes, phi=1e-4, hidden=32, group=group, precision=precision)
    f2, t2 = sg.rollout_population(bun, torch.from_numpy(genomes).cuda(), phi=1e-4, precision=precision)
    assert torch.equal(f2, fit) and torch.equal(t2, trd)
    fit, trd, act = fit.cpu().numpy(), trd.cpu().numpy(), act.cpu().numpy()
    for i in range(0, P, max(1, P // 9)):
        fo, to = orc.rollout(None, None, bz, 1e-4, 0.001, 0.0, forced_actions=act[i])
        assert fo == fit[i] and to == trd[i], (i, fo, fit[i])
    # the last individual exercises the partial group
    fo, to = orc.rollout(None, None, bz, 1e-4, 0.001, 0.0, forced_actions=act[P - 1])
    assert fo == fit[P - 1] and to == trd[P - 1]


def test_empty_bundle_and_host_entry(sg, orc):
    bundle, bz, bun, master, genomes = _setup(sg, orc, 1, 94, 6, seed=2, T=0)
    f, t = sg.rollout_population(bun, torch.from_numpy(genomes).cuda(), phi=1e-4, precision="bf16")
    assert np.array_equal(f.cpu().numpy(), np.full(6, -50.0)) and int(t.sum()) == 0      # drl_engine.py:64-65
    bundle, bz, bun, master, genomes = _setup(sg, orc, 1, 94, 37, seed=2)
    fd, td = sg.rollout_population(bun, torch.from_numpy(genomes).cuda(), phi=1e-4, precision="bf16")
    fh, th = sg.rollout_population(bun, genomes, phi=1e-4, precision="bf16")               # host pointers end to end
    assert np.array_equal(fd.cpu().numpy(), fh) and np.array_equal(td.cpu().numpy(), th)


def test_closed_loop_diverges_from_fp32_oracle_only_at_near_ties(sg, orc):
    bundle, bz, bun, master, genomes = _setup(sg, orc, 2, 92, 8, seed=5)
    fit, trd, raw, act = sg.rollout_tc_audit(bun, genomes, phi=1e-4, hidden=32)
    fit, trd, act = fit.cpu().numpy(), trd.cpu().numpy(), act.cpu().numpy()
    identical = 0
    for i in range(genomes.shape[0]):
        fo, to, tro = orc.rollout(genomes[i], None, bz, 1e-4, 0.001, 0.0, hidden=32, trace=True)
        diff = (tro["off_a"] != act[i, :, 0]) | (tro["off_b"] != act[i, :, 1])
        if not diff.any():
            identical += 1
            assert fo == fit[i] and to == trd[i]
            continue
        t = int(np.argmax(diff))
        q = np.array([tro["raw_a"][t], tro["raw_b"][t]], np.float32) * np.float32(5.0)
        margin = np.min(np.abs(np.abs(q - np.floor(q)) - 0.5))
        assert margin <= TAU_TICKS, (i, t, q)
    print("closed-loop trajectories identical to the fp32 oracle:", identical, "of", genomes.shape[0])


def test_seeded_children_match_explicit_genomes(sg, orc):
    bundle, bz, bun, master, genomes = _setup(sg, orc, 1, 93, 1, seed=3, T=100)
    kids = np.stack([orc.mutate(master, 0.05, 77, 4, 10 + i) for i in range(21)])
    f_exp, t_exp = sg.rollout_population(bun, torch.from_numpy(kids).cuda(), phi=1e-4, precision="bf16")
    f_seed, t_seed = sg.rollout_seeded(bun, torch.from_numpy(master).cuda(), count=21, sigma=0.05, seed=77,
                                       generation=4, first_index=10, phi=1e-4, precision="bf16")
    assert torch.equal(f_exp, f_seed) and torch.equal(t_exp, t_seed)


@pytest.mark.parametrize("precision", ["bf16", "tf32", "f16"])
def test_golden_arl_checkpoint_audit_set(sg, orc, precision):
    """The fixed audit set: the reference's shipped ARL agent on its own 960-bar test bundle (tests/golden,
    960/960 recorded actions).  On the (bar, inventory) pairs the recorded run visited, the tensor-core
    policy outputs are within tolerance of the fp32 oracle; a recorded offset may flip only where the
    oracle's own raw*5 is within TAU_TICKS of a rounding boundary; flips are counted and reported."""
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    c = np.load(os.path.join(here, "golden", "checkpoints.npz"))
    b = np.load(os.path.join(here, "golden", "backtest_510300.npz"))
    g = c["510300_with_adv"]
    s1, s2 = b["arl.s1_pred"], b["arl.s2_pred"]
    z1 = ((s1 - c["train_stats_s1_m"][()]) / c["train_stats_s1_s"][()]).astype(np.float32)
    z2 = ((s2 - c["train_stats_s2_m"][()]) / c["train_stats_s2_s"][()]).astype(np.float32)
    fb, fs = b["arl.fill_buy"], b["arl.fill_sell"]
    my_ask = b["arl.ask"] + b["arl.off_a"] * 0.001
    my_bid = b["arl.bid"] - b["arl.off_b"] * 0.001
    bun = sg.Bundle(z1, z2, b["arl.mid"], b["arl.ask"], b["arl.bid"],
                    np.where(fs == 1, my_ask, my_ask - 0.0005), np.where(fb == 1, my_bid, my_bid + 0.0005), 0.001)
    TAU_TICKS = TAU[precision]
    fit, trd, raw, act = sg.rollout_tc_audit(bun, g.reshape(1, -1), phi=1e-4, hidden=32, precision=precision)
    raw, act = raw.cpu().numpy()[0], act.cpu().numpy()[0]
    inv_prev = np.concatenate([[0], b["arl.inventory"][:-1]])
    worst, flips, flips_outside = 0.0, 0, 0
    for t in range(960):
        want = orc.mlp_forward(g, [z1[t], z2[t], inv_prev[t] / 2.0])
        q_o, q_k = want * np.float32(5.0), raw[t, inv_prev[t] + 2] * np.float32(5.0)
        worst = max(worst, float(np.max(np.abs(q_o - q_k))))
        margin = np.abs(np.abs(q_o - np.floor(q_o)) - 0.5)
        rec = np.array([b["arl.off_a"][t], b["arl.off_b"][t]])
        flips += int(np.sum(np.rint(q_k) != rec))
        flips_outside += int(np.sum((np.rint(q_k) != rec) & (margin > TAU_TICKS)))
    print(f"golden ARL audit set ({precision}): max |d(raw*5)| = {worst:.4g} ticks, {flips} of 1920 recorded offsets flip "
          f"(all within {TAU_TICKS} tick of a rounding boundary of the fp32 result)")
    assert worst <= TAU_TICKS and flips_outside == 0
    if precision == "tf32":
        # no recorded offset flips, so the tensor-core path walks the reference's shipped backtest EXACTLY:
        # every action, the trade count, and the fitness (same fp64 sums in the same order) are the recorded ones
        assert flips == 0
        assert np.array_equal(act[:, 0], b["arl.off_a"]) and np.array_equal(act[:, 1], b["arl.off_b"])
        assert int(trd.item()) == int(((fb == 1) | (fs == 1)).sum())
        assert fit.item() == np.cumsum(b["arl.reward"])[-1]


def test_device_ga_with_tensor_core_population(sg, orc):
    """sgmm_ga_config.precision = BF16: the population is evaluated by the tensor-core rollout, the
    argmax / tell / master update are unchanged, and the validation rollout of the best child is the
    exact fp32 kernel (bit-identical to the oracle)."""
    from sgmm_b200 import synthetic
    from sgmm_b200.engine import DeviceGA
    bundle, bz, bun, master, genomes = _setup(sg, orc, 1, 95, 1, seed=4)
    vb = synthetic.synthetic_bundle(1, first_day=96)
    stats = synthetic.train_stats_of(synthetic.synthetic_bundle(1, first_day=95))
    val = sg.Bundle.from_arrays(vb, stats, 0.001)
    vz1, vz2 = orc.normalise(vb, stats)
    ga = DeviceGA(master, None, pop_size=70, sigma=0.05, phi=1e-4, fee_rate=0.0, use_arl=False, seed=9,
                  max_generations=4, precision="bf16")
    try:
        ga.generation(bun, val)
        h = ga.history(1)
        mm, adv, best = ga.masters()
    finally:
        ga.close()
    f, t = sg.rollout_seeded(bun, torch.from_numpy(master).cuda(), count=70, sigma=0.05, seed=9, generation=0,
                             phi=1e-4, precision="bf16")
    f, t = f.cpu().numpy(), t.cpu().numpy()
    i = int(np.argmax(f))
    assert h["train_f"][0] == f[i] and h["train_trades"][0] == t[i]
    child = orc.mutate(master, 0.05, 9, 0, i)
    assert np.array_equal(mm, child)
    fo, to = orc.rollout(child, None, (vz1, vz2) + vb[2:], 1e-4, 0.001, 0.0)
    assert h["val_f"][0] == fo and h["val_trades"][0] == to
    # adversarial co-training on the tensor-core path: both masters follow the composed rollout's argmax / argmin
    adv_master = (np.random.default_rng(6).standard_normal(1250) * 0.5).astype(np.float32)
    ga = DeviceGA(master, adv_master, pop_size=50, sigma=0.05, phi=1e-4, fee_rate=0.0, use_arl=True, seed=13,
                  max_generations=2, precision="f16")
    try:
        ga.generation(bun, val)
        h = ga.history(1)
        mm, adv, best = ga.masters()
    finally:
        ga.close()
    f, t = sg.rollout_seeded(bun, torch.from_numpy(master).cuda(), count=50, sigma=0.05, seed=13, generation=0,
                             adv_master=torch.from_numpy(adv_master).cuda(), phi=1e-4, precision="f16")
    f, t = f.cpu().numpy(), t.cpu().numpy()
    i, ia = int(np.argmax(f)), int(np.argmax(-f))
    assert h["train_f"][0] == f[i] and h["train_trades"][0] == t[i]
    assert np.array_equal(mm, orc.mutate(master, 0.05, 13, 0, i))
    assert np.array_equal(adv, orc.mutate(adv_master, 0.05, 13 ^ 0x8000000000000000, 0, ia))


# ----------------------------------------------------------------------------------------------
# the adversary on the tensor-core path (BASELINE config 3): a 20-state automaton in the walker
# ----------------------------------------------------------------------------------------------
def _adv_genomes(P, seed, scale=0.7):
    return (np.random.default_rng(seed).standard_normal((P, 1250)) * scale).astype(np.float32)


@pytest.mark.parametrize("precision", ["bf16", "tf32", "f16"])
@pytest.mark.parametrize("fee", [0.0, 3e-4])
def test_adversary_env_bit_exact_given_offsets(sg, orc, fee, precision):
    """With the adversary: GIVEN the market maker's offsets the kernel took, the displacement (the oracle's own
    adv_forward on [inv/2, fill_sell_prev, fill_buy_prev]), fills, inventory, trades, rewards and fitness are
    bit-identical to the oracle's (teacher-forced replay WITH the adversary genome); policy outputs within tolerance."""
    bundle, bz, bun, master, genomes = _setup(sg, orc, 1, 97, 6, seed=21)
    adv = _adv_genomes(6, 22)
    T = bun.T
    fit, trd, raw, act = sg.rollout_tc_audit(bun, genomes, adv, phi=1e-4, fee_rate=fee, hidden=32, precision=precision)
    f2, t2 = sg.rollout_population(bun, torch.from_numpy(genomes).cuda(), torch.from_numpy(adv).cuda(), phi=1e-4, fee_rate=fee,
                                   precision=precision)
    assert torch.equal(f2, fit) and torch.equal(t2, trd)
    fit, trd, raw, act = fit.cpu().numpy(), trd.cpu().numpy(), raw.cpu().numpy(), act.cpu().numpy()
    displaced = 0
    for i in range(genomes.shape[0]):
        fo, to, tro = orc.rollout(None, adv[i], bz, 1e-4, 0.001, fee, forced_actions=act[i], trace=True)
        assert fo == fit[i] and to == trd[i], (i, fo, fit[i], to, trd[i])
        inv_before = np.concatenate([[0], tro["inventory"][:-1]])
        taken = np.rint(raw[i, np.arange(T), inv_before + 2] * np.float32(5.0)).astype(np.int32)
        assert np.array_equal(taken, act[i])
        displaced += int(np.sum((tro["adv_a"] != 0) | (tro["adv_b"] != 0)))
    assert displaced > 0, "the adversaries of this case should displace some quotes"
    worst, flips_outside = _audit(orc, bz, genomes[:2], raw[:2], T, TAU[precision])
    assert worst <= TAU[precision] and flips_outside == 0
    # the adversary matters
    f0, _ = sg.rollout_population(bun, torch.from_numpy(genomes).cuda(), phi=1e-4, fee_rate=fee, precision=precision)
    assert not np.array_equal(f0.cpu().numpy(), fit)


@pytest.mark.parametrize("T,P,group", [(1, 3, 0), (25, 150, 0), (26, 17, 16), (51, 33, 6), (240, 300, 0)])
def test_adversary_ragged_lengths_and_groups(sg, orc, T, P, group):
    precision = "f16" if (T % 2) else "tf32"
    bundle, bz, bun, master, genomes = _setup(sg, orc, 1, 98, P, seed=T, T=T)
    adv = _adv_genomes(P, 100 + T)
    fit, trd, raw, act = sg.rollout_tc_audit(bun, genomes, adv, phi=1e-4, hidden=32, group=group, precision=precision)
    fit, trd, act = fit.cpu().numpy(), trd.cpu().numpy(), act.cpu().numpy()
    for i in list(range(0, P, max(1, P // 9))) + [P - 1]:
        fo, to = orc.rollout(None, adv[i], bz, 1e-4, 0.001, 0.0, forced_actions=act[i])
        assert fo == fit[i] and to == trd[i], (i, fo, fit[i])


def test_adversary_seeded_children_and_74_float_genomes(sg, orc):
    bundle, bz, bun, master, genomes = _setup(sg, orc, 1, 99, 1, seed=3, T=120)
    adv_master = _adv_genomes(1, 7)[0]
    kids = np.stack([orc.mutate(master, 0.05, 77, 4, 10 + i) for i in range(9)])
    akids = np.stack([orc.mutate(adv_master, 0.05, 77 ^ 0x8000000000000000, 4, 10 + i) for i in range(9)])
    f_exp, t_exp = sg.rollout_population(bun, torch.from_numpy(kids).cuda(), torch.from_numpy(akids).cuda(), phi=1e-4, precision="f16")
    f_seed, t_seed = sg.rollout_seeded(bun, torch.from_numpy(master).cuda(), count=9, sigma=0.05, seed=77, generation=4, first_index=10,
                                       adv_master=torch.from_numpy(adv_master).cuda(), phi=1e-4, precision="f16")
    assert torch.equal(f_exp, f_seed) and torch.equal(t_exp, t_seed)
    # a native 74-float AdversaryPolicy genome (models/model.py:52-57 reads only the first 74 floats)
    f74, t74 = sg.rollout_population(bun, torch.from_numpy(kids).cuda(), torch.from_numpy(akids[:, :74].copy()).cuda(), phi=1e-4, precision="f16")
    assert torch.equal(f74, f_exp) and torch.equal(t74, t_exp)
    fx, tx = sg.rollout_population(bun, torch.from_numpy(kids).cuda(), torch.from_numpy(akids[:, :74].copy()).cuda(), phi=1e-4)     # exact path
    fy, ty = sg.rollout_population(bun, torch.from_numpy(kids).cuda(), torch.from_numpy(akids).cuda(), phi=1e-4)
    assert torch.equal(fx, fy) and torch.equal(tx, ty)


@pytest.mark.parametrize("out_scale,fee,use_adv", [(300.0, 0.0, False), (40000.0, 3e-4, True), (4e9, 3e-5, True)])
def test_offsets_outside_the_leg_table_take_the_literal_path(sg, orc, out_scale, fee, use_adv):
    """Offsets of hundreds to billions of ticks (both signs, clamped to +-2^30 before the displacement): fills far inside
    the threshold fall outside the 16-entry leg table of the bar and are accounted by the literal fp64 expression."""
    bundle, bz, bun, master, genomes = _setup(sg, orc, 1, 93, 9, seed=3, T=333, out_scale=out_scale)
    adv = _adv_genomes(9, 5) if use_adv else None
    fit, trd, raw, act = sg.rollout_tc_audit(bun, genomes, adv, phi=1e-4, fee_rate=fee, hidden=32, precision="tf32")
    fit, trd, act = fit.cpu().numpy(), trd.cpu().numpy(), act.cpu().numpy()
    assert np.abs(act).max() > 50 * min(out_scale, 1e6) / 300.0
    for i in range(genomes.shape[0]):
        fo, to = orc.rollout(None, adv[i] if use_adv else None, bz, 1e-4, 0.001, fee, forced_actions=act[i])
        assert fo == fit[i] and to == trd[i], (i, fo, fit[i], to, trd[i])
