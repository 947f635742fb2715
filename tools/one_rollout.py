"""One population rollout (for ncu).  args: P days reps units warps"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sgmm_b200
from sgmm_b200 import synthetic
P = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
days = int(sys.argv[2]) if len(sys.argv) > 2 else 60
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2
u = int(sys.argv[4]) if len(sys.argv) > 4 else 0
w = int(sys.argv[5]) if len(sys.argv) > 5 else 0
bundle = synthetic.synthetic_bundle(days)
bun = sgmm_b200.Bundle.from_arrays(bundle, synthetic.train_stats_of(bundle), 0.001)
_, genomes = synthetic.policy_like_genomes(P, seed=0)
g = torch.from_numpy(genomes).cuda()
for _ in range(n):
    f, t = sgmm_b200.rollout_population(bun, g, phi=1e-4, units_per_lane=u, warps_per_cta=w)
torch.cuda.synchronize()
print("ok", f.sum().item())
