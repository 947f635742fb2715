"""Sweep the exact kernel's launch geometry (units_per_lane U, warps_per_cta W) over population sizes, with and without the
adversary: the data behind the auto-tuner's buckets.  args: days"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, sgmm_b200
from sgmm_b200 import synthetic
days = int(sys.argv[1]) if len(sys.argv) > 1 else 20
bundle = synthetic.synthetic_bundle(days)
bun = sgmm_b200.Bundle.from_arrays(bundle, synthetic.train_stats_of(bundle), 0.001)
rows = []
for P in (148, 300, 592, 1024, 1500, 2048, 2500, 3000, 3552, 4096, 4144, 6000, 8192):
    _, genomes = synthetic.policy_like_genomes(P, seed=0)
    g = torch.from_numpy(genomes).cuda()
    adv = torch.from_numpy((np.random.default_rng(1).standard_normal((P, 1250)) * 0.5).astype(np.float32)).cuda()
    for use_adv in (False, True):
        best = None
        for U, W in ((0, 0), (1, 0), (2, 0), (4, 0), (2, 8), (2, 10), (2, 12), (1, 12), (1, 16)):
            try:
                run = lambda: sgmm_b200.rollout_population(bun, g, adv if use_adv else None, phi=1e-4, units_per_lane=U, warps_per_cta=W)
                run(); torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); run(); run(); e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 2
            except Exception as ex:
                continue
            r = {"P": P, "adv": use_adv, "U": U, "W": W, "ms": round(ms, 4), "G": round(P * bun.T / ms / 1e6, 2)}
            rows.append(r)
            if U and (best is None or ms < best["ms"]):
                best = r
        auto = [r for r in rows if r["P"] == P and r["adv"] == use_adv and r["U"] == 0][0]
        print(f"P={P:5d} adv={int(use_adv)} auto {auto['ms']:.3f} ms ({auto['G']} G)  best U={best['U']} W={best['W']} {best['ms']:.3f} ms ({best['G']} G)  "
              + " ".join(f"U{r['U']}W{r['W']}={r['ms']:.3f}" for r in rows if r["P"] == P and r["adv"] == use_adv and r["U"]), flush=True)
json.dump(rows, open("gpurun_out/exact_sweep.json", "w"))
