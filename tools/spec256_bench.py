"""Time the H=256 tensor-core rollout (BASELINE config 4 shape).  args: P days reps [explicit]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sgmm_b200
from sgmm_b200 import synthetic
P = int(sys.argv[1]) if len(sys.argv) > 1 else 592
days = int(sys.argv[2]) if len(sys.argv) > 2 else 10
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
explicit = len(sys.argv) > 4
bundle = synthetic.synthetic_bundle(days)
bun = sgmm_b200.Bundle.from_arrays(bundle, synthetic.train_stats_of(bundle), 0.001)
master, genomes = synthetic.policy_like_genomes(P if explicit else 1, hidden=256, seed=0)
m = torch.from_numpy(master).cuda()
g = torch.from_numpy(genomes).cuda() if explicit else None
def run():
    if explicit:
        return sgmm_b200.rollout_population(bun, g, phi=1e-4, fee_rate=3e-4, hidden=256)
    return sgmm_b200.rollout_seeded(bun, m, count=P, sigma=0.05, seed=1, generation=0, phi=1e-4, fee_rate=3e-4, hidden=256)
for _ in range(2):
    f, t = run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    f, t = run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
steps = P * bun.T
peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except Exception:
    pass
peak = peaks.get("bf16_tflops_sustained", 1400.0)
alg = steps * 133632 / ms / 1e9
print(json.dumps({"kernel": "spec256_kernel", "P": P, "T": bun.T, "genomes": "explicit" if explicit else "seeded (Philox in-kernel)",
                  "ms": ms, "env_steps_per_s": steps / ms * 1e3, "algorithmic_tflops": alg,
                  "executed_hidden_tflops": steps * 5.12 * 131072 / ms / 1e9, "tensor_peak_tflops_sustained": peak,
                  "roofline_frac_algorithmic": alg / peak, "checksum": f.sum().item()}))
