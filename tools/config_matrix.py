"""Time the rollout across the BASELINE configs' feature matrix (adversary / fee / seeded children)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sgmm_b200
from sgmm_b200 import synthetic

def timed(fn, reps=3):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

bundle = synthetic.synthetic_bundle(60)
bun = sgmm_b200.Bundle.from_arrays(bundle, synthetic.train_stats_of(bundle), 0.001)
T = bun.T
rows = []
for P in (50, 2048, 4096, 8192, 16384):
    master, genomes = synthetic.policy_like_genomes(P, seed=0)
    g = torch.from_numpy(genomes).cuda()
    m = torch.from_numpy(master).cuda()
    adv = torch.from_numpy((np.random.default_rng(1).standard_normal((P, 1250)) * 0.7).astype(np.float32)).cuda()
    advm = adv[0].contiguous()
    cases = {
        "explicit": lambda: sgmm_b200.rollout_population(bun, g, phi=1e-4),
        "explicit+fee": lambda: sgmm_b200.rollout_population(bun, g, phi=1e-4, fee_rate=3e-4),
        "explicit+adversary": lambda: sgmm_b200.rollout_population(bun, g, adv, phi=1e-4),
        "seeded": lambda: sgmm_b200.rollout_seeded(bun, m, count=P, sigma=0.05, seed=1, generation=0, phi=1e-4),
        "seeded+adversary": lambda: sgmm_b200.rollout_seeded(bun, m, count=P, sigma=0.05, seed=1, generation=0, adv_master=advm, phi=1e-4),
        "tensor tf32 explicit": lambda: sgmm_b200.rollout_population(bun, g, phi=1e-4, precision="tf32"),
        "tensor tf32 explicit+fee": lambda: sgmm_b200.rollout_population(bun, g, phi=1e-4, fee_rate=3e-4, precision="tf32"),
        "tensor tf32 seeded": lambda: sgmm_b200.rollout_seeded(bun, m, count=P, sigma=0.05, seed=1, generation=0, phi=1e-4, precision="tf32"),
        "tensor f16 explicit": lambda: sgmm_b200.rollout_population(bun, g, phi=1e-4, precision="f16"),
        "tensor f16 seeded": lambda: sgmm_b200.rollout_seeded(bun, m, count=P, sigma=0.05, seed=1, generation=0, phi=1e-4, precision="f16"),
        "tensor bf16 explicit": lambda: sgmm_b200.rollout_population(bun, g, phi=1e-4, precision="bf16"),
        "tensor bf16 seeded": lambda: sgmm_b200.rollout_seeded(bun, m, count=P, sigma=0.05, seed=1, generation=0, phi=1e-4, precision="bf16"),
    }
    for name, fn in cases.items():
        ms = timed(fn)
        rows.append({"P": P, "T": T, "case": name, "ms": round(ms, 3), "G_env_steps_per_s": round(P * T / ms / 1e6, 3)})
        print(rows[-1], flush=True)
# BASELINE config 4 at its full size: H = 256, with fee, population 16384 (seeded children), 120 synthetic days
bundle4 = synthetic.synthetic_bundle(120)
bun4 = sgmm_b200.Bundle.from_arrays(bundle4, synthetic.train_stats_of(bundle4), 0.001)
m256, _ = synthetic.policy_like_genomes(1, hidden=256, seed=0)
m256 = torch.from_numpy(m256).cuda()
for P in (16384,):
    ms = timed(lambda: sgmm_b200.rollout_seeded(bun4, m256, count=P, sigma=0.05, seed=1, generation=0, phi=1e-4, fee_rate=3e-4, hidden=256), reps=2)
    rows.append({"P": P, "T": bun4.T, "case": "config 4: H=256 tensor cores, fee, seeded", "ms": round(ms, 3), "G_env_steps_per_s": round(P * bun4.T / ms / 1e6, 3)})
    print(rows[-1], flush=True)
json.dump(rows, open("gpurun_out/config_matrix_r1b.json", "w"), indent=1)
