"""A few exact (F32) population rollouts of the bench workload: for launch lists / ncu captures.  args: P days reps"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, sgmm_b200
from sgmm_b200 import synthetic
P = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
days = int(sys.argv[2]) if len(sys.argv) > 2 else 60
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
bundle = synthetic.synthetic_bundle(days)
bun = sgmm_b200.Bundle.from_arrays(bundle, synthetic.train_stats_of(bundle), 0.001)
master, genomes = synthetic.policy_like_genomes(P, hidden=32, seed=0, out_scale=6.0, out_bias=(0.1, 0.1))
g = torch.from_numpy(genomes).cuda()
for _ in range(reps):
    f, t = sgmm_b200.rollout_population(bun, g, phi=1e-4)
torch.cuda.synchronize()
print("fitness mean", f.mean().item(), "trades mean", t.double().mean().item())
