"""One tensor-core H=32 rollout (for ncu).  args: P days reps group fee precision"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sgmm_b200
from sgmm_b200 import synthetic
P = int(sys.argv[1]) if len(sys.argv) > 1 else 2072
days = int(sys.argv[2]) if len(sys.argv) > 2 else 10
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2
grp = int(sys.argv[4]) if len(sys.argv) > 4 else 0
fee = float(sys.argv[5]) if len(sys.argv) > 5 else 0.0
prec = sys.argv[6] if len(sys.argv) > 6 else "bf16"
bundle = synthetic.synthetic_bundle(days)
bun = sgmm_b200.Bundle.from_arrays(bundle, synthetic.train_stats_of(bundle), 0.001)
_, genomes = synthetic.policy_like_genomes(P, seed=0, out_scale=6.0, out_bias=(0.1, 0.1))
g = torch.from_numpy(genomes).cuda()
for _ in range(n):
    f, t = sgmm_b200.rollout_population(bun, g, phi=1e-4, fee_rate=fee, precision=prec, units_per_lane=grp)
torch.cuda.synchronize()
print("ok", f.sum().item())
