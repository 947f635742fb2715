"""Time the rollout kernel variants (units per lane x warps per CTA) at BASELINE config 2 sizes."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sgmm_b200
from sgmm_b200 import synthetic

P = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
days = int(sys.argv[2]) if len(sys.argv) > 2 else 60
combos = [(4, 0), (4, 4), (2, 0), (2, 8), (1, 0), (1, 8)]
if len(sys.argv) > 3:
    combos = [tuple(int(x) for x in c.split(":")) for c in sys.argv[3].split(",")]
bundle = synthetic.synthetic_bundle(days)
stats = synthetic.train_stats_of(bundle)
bun = sgmm_b200.Bundle.from_arrays(bundle, stats, 0.001)
_, genomes = synthetic.policy_like_genomes(P, seed=0)
g = torch.from_numpy(genomes).cuda()
for (u, w) in combos:
    for _ in range(2):
        f, t = sgmm_b200.rollout_population(bun, g, phi=1e-4, units_per_lane=u, warps_per_cta=w)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        f, t = sgmm_b200.rollout_population(bun, g, phi=1e-4, units_per_lane=u, warps_per_cta=w)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"U={u} W={w:2d}  {ms:8.3f} ms  {P * bun.T / ms / 1e6:8.2f} G env-steps/s  checksum {f.sum().item():.6f}", flush=True)
