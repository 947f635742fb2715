"""Time the H=32 tensor-core rollout against the exact fp32 kernel (BASELINE config 2 shape), with and without the adversary.
args: P days reps [group]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sgmm_b200
from sgmm_b200 import synthetic
P = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
days = int(sys.argv[2]) if len(sys.argv) > 2 else 60
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
group = int(sys.argv[4]) if len(sys.argv) > 4 else 0
bundle = synthetic.synthetic_bundle(days)
bun = sgmm_b200.Bundle.from_arrays(bundle, synthetic.train_stats_of(bundle), 0.001)
master, genomes = synthetic.policy_like_genomes(P, hidden=32, seed=0, out_scale=1.0)
g = torch.from_numpy(genomes).cuda()
adv = torch.from_numpy((np.random.default_rng(3).standard_normal((P, 1250)) * 0.5).astype(np.float32)).cuda()
out = {"P": P, "T": bun.T}
for use_adv in (False, True):
    for prec in ("f16", "bf16", "tf32", "f32"):
        def run():
            return sgmm_b200.rollout_population(bun, g, adv if use_adv else None, phi=1e-4, precision=prec,
                                                units_per_lane=group if prec != "f32" else 0)
        for _ in range(2):
            f, t = run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            f, t = run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        out[prec + ("+adv" if use_adv else "")] = {"ms": round(ms, 4), "G_env_steps_per_s": round(P * bun.T / ms / 1e6, 3),
                                                   "fitness_mean": f.mean().item(), "trades_mean": t.double().mean().item()}
        print(prec, "adv" if use_adv else "", out[prec + ("+adv" if use_adv else "")], flush=True)
print(json.dumps(out))
