"""One exact adversarial population rollout (BASELINE config 3 shape, for ncu).  args: P days"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, sgmm_b200
from sgmm_b200 import synthetic
P = int(sys.argv[1]); days = int(sys.argv[2])
bundle = synthetic.synthetic_bundle(days)
bun = sgmm_b200.Bundle.from_arrays(bundle, synthetic.train_stats_of(bundle), 0.001)
_, genomes = synthetic.policy_like_genomes(P, seed=0)
g = torch.from_numpy(genomes).cuda()
adv = torch.from_numpy((np.random.default_rng(3).standard_normal((P, 1250)) * 0.5).astype(np.float32)).cuda()
for _ in range(2):
    f, t = sgmm_b200.rollout_population(bun, g, adv, phi=1e-4)
torch.cuda.synchronize()
print("ok", f.sum().item())
