"""Smallest invocation of every kernel (for compute-sanitizer)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sgmm_b200
from sgmm_b200 import synthetic
from sgmm_b200.engine import DeviceGA
b = tuple(a[:150] for a in synthetic.synthetic_bundle(1, first_day=5))
st = synthetic.train_stats_of(b)
bun = sgmm_b200.Bundle.from_arrays(b, st, 0.001)
m, g = synthetic.policy_like_genomes(9, seed=1, out_scale=6.0)
adv = np.random.default_rng(0).standard_normal((9, 1250)).astype(np.float32)
for u in (1, 2, 4):
    f, t = sgmm_b200.rollout_population(bun, torch.from_numpy(g).cuda(), torch.from_numpy(adv).cuda(), phi=1e-4, fee_rate=3e-5, units_per_lane=u)
f, t = sgmm_b200.rollout_seeded(bun, torch.from_numpy(m).cuda(), count=7, sigma=0.05, seed=3, generation=1, phi=1e-4)
sgmm_b200.rollout_trace(bun, g[0], adv[0], phi=1e-4)
sgmm_b200.rollout_table(bun, sgmm_b200.FOICPolicy(0, 0).table(b), phi=1e-4)
ga = DeviceGA(m, adv[0], pop_size=10, sigma=0.05, phi=1e-4, fee_rate=0.0, use_arl=True, seed=1, max_generations=2)
ga.generation(bun, bun); ga.generation(bun, bun); ga.history(2); ga.close()
_, g256 = synthetic.policy_like_genomes(2, hidden=256, seed=2)
b2 = tuple(a[:30] for a in b)
bun2 = sgmm_b200.Bundle.from_arrays(b2, st, 0.001)
f2 = sgmm_b200.rollout_spec256_audit(bun2, g256, phi=1e-4)
torch.cuda.synchronize()
print("sanitize_small ok", f.sum().item(), f2[0].sum().item())
