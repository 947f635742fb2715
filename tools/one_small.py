"""One small-population exact rollout (for ncu).  args: P days [adv]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, sgmm_b200
from sgmm_b200 import synthetic
P = int(sys.argv[1]); days = int(sys.argv[2]); use_adv = len(sys.argv) > 3
bundle = synthetic.synthetic_bundle(days)
bun = sgmm_b200.Bundle.from_arrays(bundle, synthetic.train_stats_of(bundle), 0.001)
m, _ = synthetic.policy_like_genomes(1, seed=0)
md = torch.from_numpy(m).cuda()
am = torch.from_numpy((np.random.default_rng(2).standard_normal(1250) * 0.5).astype(np.float32)).cuda() if use_adv else None
for _ in range(2):
    f, t = sgmm_b200.rollout_seeded(bun, md, count=P, sigma=0.05, seed=1, generation=0, adv_master=am, phi=1e-4)
torch.cuda.synchronize()
print("ok", f.sum().item())
