"""Soak: many rollouts of random shapes through every path (hang / barrier-phase regressions show up as a timeout or a mismatch
between paths).  args: iterations seed"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, sgmm_b200
from sgmm_b200 import synthetic
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
base = synthetic.synthetic_bundle(6, first_day=700)
stats = synthetic.train_stats_of(base)
t0 = time.time()
for it in range(iters):
    T = int(rng.choice([1, 7, 24, 25, 26, 49, 50, 51, 127, 128, 129, 300, 777, 1440]))
    P = int(rng.choice([1, 2, 3, 13, 14, 15, 16, 17, 27, 28, 29, 64, 147, 148, 149, 297, 300, 600]))
    use_adv = bool(rng.integers(2)); fee = float(rng.choice([0.0, 3e-4]))
    bundle = tuple(a[:T] for a in base)
    bun = sgmm_b200.Bundle.from_arrays(bundle, stats, 0.001)
    _, g = synthetic.policy_like_genomes(P, seed=it, out_scale=float(rng.choice([1.0, 6.0])), out_bias=(0.1, 0.1))
    gd = torch.from_numpy(g).cuda()
    ad = torch.from_numpy((rng.standard_normal((P, 1250)) * 0.6).astype(np.float32)).cuda() if use_adv else None
    f0, t0_ = sgmm_b200.rollout_population(bun, gd, ad, phi=1e-4, fee_rate=fee)                      # auto route
    f1, t1 = sgmm_b200.rollout_population(bun, gd, ad, phi=1e-4, fee_rate=fee, units_per_lane=4)     # sequential kernel
    assert torch.equal(f0, f1) and torch.equal(t0_, t1), (it, T, P, use_adv, fee)
    for prec in ("f16", "tf32", "bf16"):
        fa, ta = sgmm_b200.rollout_population(bun, gd, ad, phi=1e-4, fee_rate=fee, precision=prec)
        fb, tb = sgmm_b200.rollout_population(bun, gd, ad, phi=1e-4, fee_rate=fee, precision=prec)
        assert torch.equal(fa, fb) and torch.equal(ta, tb), ("tensor path not deterministic", it, T, P, use_adv, prec)
        assert torch.isfinite(fa).all()
    torch.cuda.synchronize()
    bun.close()
print(f"soak ok: {iters} random shapes x 5 paths in {time.time() - t0:.1f} s")
